"""where does NCCL's INFO log go?  (bench.py: the log must reach stderr, stdout stays the one JSON line)"""
import os, sys
mode = sys.argv[1]
if mode == "early_file":
    os.environ.setdefault("NCCL_DEBUG", "INFO"); os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
elif mode == "early":
    os.environ.setdefault("NCCL_DEBUG", "INFO")
import torch, torch.distributed as dist
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = torch.ones(1, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
if dist.get_rank() == 0: print('{"ok": %d}' % int(t.item()))
dist.destroy_process_group()
