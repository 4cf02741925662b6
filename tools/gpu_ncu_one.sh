#!/bin/bash
# ncu --set full of ONE launch of one kernel (base name regex).  Usage: gpurun -- 'bash tools/gpu_ncu_one.sh TAG KERNEL SKIP [bench args]'
TAG=$1; KRE=$2; SKIP=${3:-4}; shift 3
O=gpurun_out
mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 -o $O/k_$TAG -f \
  python bench.py --seqs 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $O/ncu_k_$TAG.log 2>&1
tail -1 $O/ncu_k_$TAG.log
