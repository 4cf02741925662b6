#!/bin/bash
# Round-2 final GPU visit: parity tests, smoke, bench lines (default, S=1, 4K, rectify, reference arms), ncu launch list
# with DRAM bytes and a --set full capture of three steady-state steps.
TAG=${1:-r2s}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
grep "GPU vs\|passed\|failed\|rc=" $O/pytest_gpu_$TAG.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -1 $O/smoke_$TAG.log
timeout 900 python bench.py > $O/bench_${TAG}_s64.json 2> $O/bench_${TAG}_s64.err; tail -2 $O/bench_${TAG}_s64.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err
timeout 300 python bench.py --seqs 1 --steps 50 --no-cpu-baseline --no-extras > $O/bench_${TAG}_s1.json 2> $O/bench_${TAG}_s1.err
for S in 8 16 32; do timeout 300 python bench.py --seqs $S --steps 40 --no-cpu-baseline --no-extras > $O/bench_${TAG}_s$S.json 2> $O/bench_${TAG}_s$S.err; done
timeout 600 python bench.py --workload 4k --steps 10 > $O/bench_${TAG}_4k.json 2> $O/bench_${TAG}_4k.err
timeout 300 python bench.py --workload rectify > $O/bench_${TAG}_rectify.json 2> $O/bench_${TAG}_rectify.err
[ -n "$NO_NCU" ] || timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -c 300 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --seqs 64 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_launches_$TAG.log 2>&1
[ -n "$NO_NCU" ] || timeout 900 ncu --set full --clock-control none -k regex:"k_" -s 78 -c 19 -o $O/full_$TAG -f \
  python bench.py --seqs 64 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_full_$TAG.log 2>&1
python - <<PY
import json
d=json.load(open('$O/bench_${TAG}_s64.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),round(d['e2e']['frac_of_device_rate'],3),'f32',round(d['e2e']['f32_ingest']['value']))
print({k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})
print('kf',d['keyframe_fraction'],'pose',d['pose_error_vs_gt'],'lat',d.get('latency_s1_ms'),'4k',d['config5_4k']['value'],'cpu',d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
r=json.load(open('$O/bench_${TAG}_reference.json')); print('reference arm', r['value'], r['cpu_baseline']['cores'], r['pose_error_vs_gt'])
for S in (1,8,16,32):
    x=json.load(open('$O/bench_${TAG}_s%d.json'%S)); print('S',S,round(x['value']),round(x['ms_per_step'],3),'e2e',round(x['e2e']['value']))
x=json.load(open('$O/bench_${TAG}_4k.json')); print('4k',x['value'],x['ms_per_step'],x['roofline']['frac'],x['e2e']['value'],x['cpu_baseline'])
x=json.load(open('$O/bench_${TAG}_rectify.json')); print('rectify',x['value'],x['roofline']['frac'])
PY
ls -la $O | tail -5
