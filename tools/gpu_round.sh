#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), ncu launch list and a --set full capture.
# Usage (from the container): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -5 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -2 $O/smoke_$TAG.log
timeout 600 python bench.py > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; tail -c 600 $O/bench_${TAG}_s64.json
timeout 300 python bench.py --seqs 1 --steps 50 --no-cpu-baseline > $O/bench_${TAG}_s1.json 2> $O/bench_${TAG}_s1.err
timeout 600 python bench.py --workload 4k --steps 10 > $O/bench_${TAG}_4k.json 2> $O/bench_${TAG}_4k.err; tail -c 1200 $O/bench_${TAG}_4k.json; tail -3 $O/bench_${TAG}_4k.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err; tail -c 400 $O/bench_${TAG}_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 400 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --seqs 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_" -s 84 -c 26 -o $O/full_$TAG -f \
  python bench.py --seqs 64 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -c 80 --csv --log-file $O/launches_${TAG}_4k.csv \
  python bench.py --workload 4k --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches_${TAG}_4k.log 2>&1
timeout 300 python bench.py --workload rectify > $O/bench_${TAG}_rectify.json 2> $O/bench_${TAG}_rectify.err; tail -c 700 $O/bench_${TAG}_rectify.json
timeout 300 python bench.py --workload rectify --impl reference --steps 3 > $O/bench_${TAG}_rectify_reference.json 2> $O/bench_${TAG}_rectify_reference.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_remap" -s 4 -c 2 -o $O/full_${TAG}_rectify -f \
  python bench.py --workload rectify --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_${TAG}_rectify.log 2>&1
ls -la $O
