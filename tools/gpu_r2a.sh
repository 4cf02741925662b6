#!/bin/bash
# Round-2 first GPU visit: parity tests on the street scene, smoke, bench (default + S sweep), ncu launch list.
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -15 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -2 $O/smoke_$TAG.log
timeout 900 python bench.py > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; tail -c 3000 $O/bench_${TAG}_s64.json; tail -3 $O/bench_${TAG}_s64.err
for S in 8 16 32; do
  timeout 300 python bench.py --seqs $S --steps 40 --no-cpu-baseline --no-e2e --no-extras > $O/bench_${TAG}_s$S.json 2> $O/bench_${TAG}_s$S.err
  python -c "import json;d=json.load(open('$O/bench_${TAG}_s$S.json'));print($S, d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
done
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err; tail -c 600 $O/bench_${TAG}_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 400 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --seqs 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_launches_$TAG.log 2>&1
ls -la $O
