#!/bin/bash
# tests of the changed kernels, LM/refine cycle split, ncu --set full source-level captures of the three depth kernels
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_depth.py tests/test_gpu_lm.py tests/test_gpu_tracker.py tests/test_gpu_long.py tests/test_gpu_ref.py -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -8 $O/pytest_gpu_$TAG.log
for S in 1 8; do echo "== S=$S"; python tools/lm_cycles.py $S 2>&1 | tail -4; done
for K in k_disparity k_select_warp k_refine; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 1 -o $O/k_${TAG}_$K -f \
    python bench.py --seqs 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_k_${TAG}_$K.log 2>&1
  tail -1 $O/ncu_k_${TAG}_$K.log | cut -c 1-200
done
ls -la $O/*.ncu-rep
