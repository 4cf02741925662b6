#!/bin/bash
O=gpurun_out
B="python bench.py --steps 40 --no-cpu-baseline --no-e2e --no-extras"
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
cp odometry_b200/libodo_b200.so /tmp/lib512.so
ODO_EXTRA_NVCC="-DODO_LM_THREADS=256" python odometry_b200/build.py --force > /dev/null 2>&1 && echo built256
for S in 64 32 8 1; do $B --seqs $S > $O/t.json 2>/dev/null; show $O/t.json "LM256 S$S"; done
ODO_LM_CTAS_PER_SEQ=2 $B --seqs 64 > $O/t.json 2>/dev/null; show $O/t.json "LM256 S64 G=2"
timeout 300 python -m pytest tests/test_gpu_lm.py -m gpu -x -q 2>&1 | tail -1
cp /tmp/lib512.so odometry_b200/libodo_b200.so
