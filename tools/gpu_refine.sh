#!/bin/bash
# parity tests + default bench + refine cycle split + a 128-sequence bench (global-memory fallback of the refinement state)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('S=64', round(d['value']), d['ms_per_step'], {k: round(x,3) for k,x in d['roofline']['stage_ms_per_step'].items()}, d['all_status_ok'])"
timeout 300 python bench.py --seqs 128 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('S=128', round(d['value']), d['ms_per_step'], {k: round(x,3) for k,x in d['roofline']['stage_ms_per_step'].items()}, d['all_status_ok'])"
timeout 200 python tools/lm_cycles.py 64 2>&1 | grep refine | tail -2
