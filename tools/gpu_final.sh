#!/bin/bash
# last visit of a round: the whole GPU suite, smoke, and the two KITTI bench lines at HEAD
TAG=${1:-final}
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log; tail -3 $O/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; head -c 300 $O/bench_${TAG}_s64.json; echo
timeout 200 python bench.py --seqs 1 --steps 50 --no-cpu-baseline > $O/bench_${TAG}_s1.json 2> $O/bench_${TAG}_s1.err; head -c 250 $O/bench_${TAG}_s1.json; echo
