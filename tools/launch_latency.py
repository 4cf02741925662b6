"""Kernel-boundary cost on this GPU: the time per node of a CUDA graph of N dependent tiny kernels (what a graph-captured
LM iteration -- evaluate, reduce, solve as three kernels -- would pay per boundary), next to the same N kernels launched
eagerly on one stream.  Writes profiles/launch_latency_r2.json.  ("Why a persistent kernel".)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 600
x = torch.zeros(32, device="cuda")
s = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(s):
    for _ in range(10):
        x.add_(1.0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(N):
            x.add_(1.0)          # one tiny dependent kernel per node
    torch.cuda.synchronize()
    res = {}
    for name, fn in (("graph_replay", g.replay), ("eager_stream", lambda: [x.add_(1.0) for _ in range(N)])):
        best = 1e9
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            fn()
            e1.record(s)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / N)
        res[name + "_us_per_kernel"] = best
res["kernels"] = N
res["note"] = "dependent 1-CTA kernels; time per kernel includes its (sub-microsecond) execution"
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "profiles", "launch_latency_r2.json"), "w"), indent=1)
