#!/bin/bash
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_long.py tests/test_gpu_tracker.py -m gpu -x -q -s > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep "GPU vs\|passed\|failed" $O/pytest_gpu_$TAG.log | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; tail -3 $O/bench_${TAG}_s64.err
python - <<PY
import json
d=json.load(open('$O/bench_${TAG}_s64.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),round(d['e2e']['frac_of_device_rate'],3),'f32',round(d['e2e']['f32_ingest']['value']), 'probe', d['e2e']['h2d_gbs_probe_this_rank'])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_lm_solve" -s 2 -c 1 -o $O/k_${TAG}_lm -f \
    python bench.py --seqs 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_k_${TAG}_lm.log 2>&1
tail -1 $O/ncu_k_${TAG}_lm.log | cut -c 1-150
