#!/bin/bash
# Shortest GPU-box visit: the parity tests and one default bench line.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_tests.sh TAG'
TAG=${1:-t}
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -15 $O/pytest_gpu_$TAG.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; head -c 400 $O/bench_${TAG}_s64.json
