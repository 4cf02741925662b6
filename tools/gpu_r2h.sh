#!/bin/bash
TAG=${1:-r2h}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_lm.py tests/test_gpu_groups.py tests/test_gpu_long.py tests/test_gpu_tracker.py tests/test_gpu_depth.py -m gpu -x -q -s > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep "GPU vs\|passed\|failed\|Error" $O/pytest_gpu_$TAG.log | tail -6
python tools/lm_cycles.py 1 2>&1 | grep "^frame" | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-extras > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; tail -3 $O/bench_${TAG}_s64.err
python - <<PY
import json
d=json.load(open('$O/bench_${TAG}_s64.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),round(d['e2e']['frac_of_device_rate'],3),'f32',round(d['e2e']['f32_ingest']['value']), 'probe', round(d['e2e']['h2d_gbs_probe_this_rank'],1))
print({k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})
PY
