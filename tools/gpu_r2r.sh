#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_long.py tests/test_gpu_groups.py tests/test_gpu_camera.py tests/test_shard.py tests/test_gpu_ref.py -m gpu -x -q -s 2>&1 | grep "GPU vs\|passed\|failed\|Error" | tail -4
B="python bench.py --steps 40 --no-cpu-baseline --no-extras"
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
for S in 64 32 16 8 1; do $B --seqs $S > $O/t.json 2>/dev/null; show $O/t.json "S$S"; done
