#!/bin/bash
# ncu --set full of ONE kernel (regex) with an environment switch off and on.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_ncu_kernel.sh TAG KERNEL_REGEX ENVVAR [SKIP]'
TAG=$1; KRE=$2; VAR=${3:-ODO_NONE}; SKIP=${4:-4}
O=gpurun_out
mkdir -p $O
for v in 0 1; do
  if [ $v = 1 ]; then export $VAR=1; fi
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 -o $O/k_${TAG}_$v -f \
    python bench.py --seqs 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_k_${TAG}_$v.log 2>&1
  tail -2 $O/ncu_k_${TAG}_$v.log
done
ls -la $O/*.ncu-rep
