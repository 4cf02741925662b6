#!/bin/bash
# ncu --set full of one k_lm_solve launch of the 4K dense workload.  Usage: gpurun -- 'bash tools/gpu_ncu_4k.sh TAG'
TAG=${1:-4k}
O=gpurun_out
mkdir -p $O
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_lm_solve" -s 3 -c 1 -o $O/k_lm4k_$TAG -f \
  python bench.py --workload 4k --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_k_lm4k_$TAG.log 2>&1
tail -1 $O/ncu_k_lm4k_$TAG.log
