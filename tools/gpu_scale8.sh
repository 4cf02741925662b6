#!/bin/bash
# 8-GPU runs: weak scaling (64 sequences per GPU) with the e2e leg, and config 4 as written (64 sequences in total)
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo_$TAG.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $O/scale8_weak_$TAG.json 2> $O/scale8_weak_$TAG.err
python - <<PY
import json
for l in open('$O/scale8_weak_$TAG.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print('WEAK N=8 value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(e['value']),round(e['frac_of_device_rate'],3),'h2d/rank',round(e['h2d_gbs_achieved_per_rank'],1),'probe',round(e['h2d_gbs_probe_this_rank'],1),e['host_placement'],'f32',round(e['f32_ingest']['value']))
PY
timeout 600 $TR bench.py --gpus 8 --seqs-total 64 --steps 40 --warmup 3 --no-cpu-baseline --no-extras > $O/scale8_strong_$TAG.json 2> $O/scale8_strong_$TAG.err
python - <<PY
import json
for l in open('$O/scale8_strong_$TAG.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print('STRONG N=8 (8 seq/GPU) value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(e['value']),d['scaling'],{k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})
PY
grep -h "NCCL INFO.*nranks\|comm 0x" $O/scale8_weak_$TAG.err | head -3
tail -2 $O/scale8_weak_$TAG.err | cut -c 1-200
