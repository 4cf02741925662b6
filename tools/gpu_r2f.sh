#!/bin/bash
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -6 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -2 $O/smoke_$TAG.log
timeout 900 python bench.py > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; tail -3 $O/bench_${TAG}_s64.err
python - <<PY
import json
d=json.load(open('$O/bench_${TAG}_s64.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),round(d['e2e']['frac_of_device_rate'],3),'f32',round(d['e2e']['f32_ingest']['value']))
print({k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})
print('kf',d['keyframe_fraction'],'pose',d['pose_error_vs_gt']['median'],d['pose_error_vs_gt']['max'],'lat',d.get('latency_s1_ms'),'4k',d['config5_4k']['value'],'cpu',d['cpu_baseline']['value'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -c 300 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --seqs 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_launches_$TAG.log 2>&1
tail -2 $O/ncu_launches_$TAG.log | cut -c 1-200
