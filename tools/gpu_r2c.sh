#!/bin/bash
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_depth.py tests/test_gpu_groups.py tests/test_gpu_tracker.py tests/test_gpu_long.py -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -8 $O/pytest_gpu_$TAG.log
B="python bench.py --steps 40 --no-cpu-baseline --no-e2e --no-extras"
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
$B --seqs 64 > $O/b_${TAG}_64.json 2>/dev/null; show $O/b_${TAG}_64.json "S64 new"
ODO_SELECT_V1=1 $B --seqs 64 > $O/b_${TAG}_64v1.json 2>/dev/null; show $O/b_${TAG}_64v1.json "S64 selectV1"
for S in 8 16 32; do
  $B --seqs $S > $O/b_${TAG}_$S.json 2>/dev/null; show $O/b_${TAG}_$S.json "S$S new"
  ODO_REFINE_WHOLE_DEVICE=1 $B --seqs $S > $O/b_${TAG}_${S}w.json 2>/dev/null; show $O/b_${TAG}_${S}w.json "S$S refine-whole-device"
done
$B --seqs 1 > $O/b_${TAG}_1.json 2>/dev/null; show $O/b_${TAG}_1.json "S1"
