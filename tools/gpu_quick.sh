#!/bin/bash
# Short GPU-box visit: parity tests, smoke, the KITTI bench line, the 4K (config 5) bench line and its reference arm.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_quick.sh TAG'
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -5 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -2 $O/smoke_$TAG.log
timeout 600 python bench.py > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; tail -c 300 $O/bench_${TAG}_s64.json
timeout 600 python bench.py --workload 4k --steps 10 > $O/bench_${TAG}_4k.json 2> $O/bench_${TAG}_4k.err; tail -c 1500 $O/bench_${TAG}_4k.json; tail -3 $O/bench_${TAG}_4k.err
timeout 300 python bench.py --streams 2 --no-cpu-baseline --no-e2e > $O/bench_${TAG}_s64x2.json 2> $O/bench_${TAG}_s64x2.err; head -c 400 $O/bench_${TAG}_s64x2.json
timeout 400 python bench.py --impl reference --workload 4k --steps 1 > $O/bench_${TAG}_4k_reference.json 2> $O/bench_${TAG}_4k_reference.err; head -c 300 $O/bench_${TAG}_4k_reference.json
ls -la $O
