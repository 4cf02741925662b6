"""Profiling aid: cycle split of k_lm_solve (evaluation / reduction+barrier / serial LM step)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from odometry_b200 import capi
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
fr = bench.render_sequences(list(range(S)), 6, "cuda").to(torch.float32)
ctx = capi.Context(0)
trk = capi.Tracker(ctx, 1241, 376, 4, S)
L = fr[:, :, 0].contiguous(); R = fr[:, :, 1].contiguous()
torch.cuda.synchronize()   # device-resident frames must be complete before odo_tracker_step_dev
for f in range(6):
    trk.step_dev(L[f].data_ptr(), R[f].data_ptr())
    if f:
        out = (C.c_longlong * 8)(); it = (C.c_int * 8)()
        capi.lib().odo_tracker_lm_cycles(trk.h, out, it)
        n = sum(it[:4])
        rc_ = (C.c_longlong * 8)(); capi.lib().odo_tracker_refine_cycles(trk.h, rc_)
        print(f"   refine: first {rc_[0]} pass/iter {rc_[1]/max(1,rc_[5]-0):.0f} reduce/iter {rc_[2]/max(1,rc_[5]+1):.0f} loop {rc_[3]} final {rc_[4]} iters {rc_[5]} G {rc_[6]} nsel {rc_[7]}  stage {trk.stage_ms()['depth_refine']:.3f} ms")
        print(f"frame {f}: iters {list(it)[:4]} total {out[3]} cyc = {out[3]/1.965e3:.1f} us | per iter: eval {out[0]/n:.0f} reduce {out[1]/n:.0f} step {out[2]/n:.0f} cyc [decide {out[4]/n:.0f} qr {out[5]/n:.0f} se3 {out[6]/n:.0f}]", trk.stage_ms()["lm_solve"])
