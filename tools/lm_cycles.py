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
for f in range(6):
    trk.step_dev(L[f].data_ptr(), R[f].data_ptr())
    if f:
        out = (C.c_longlong * 8)(); it = (C.c_int * 8)()
        capi.lib().odo_tracker_lm_cycles(trk.h, out, it)
        n = sum(it[:4])
        print(f"frame {f}: iters {list(it)[:4]} total {out[3]} cyc = {out[3]/1.965e3:.1f} us | per iter: eval {out[0]/n:.0f} reduce {out[1]/n:.0f} step {out[2]/n:.0f} cyc [decide {out[4]/n:.0f} qr {out[5]/n:.0f} se3 {out[6]/n:.0f}]", trk.stage_ms()["lm_solve"])
