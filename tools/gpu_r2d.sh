#!/bin/bash
# tests after the dense-LM / select / group changes, 4K A/B, compute-sanitizer passes over the persistent kernels
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_lm.py tests/test_gpu_groups.py tests/test_gpu_depth.py tests/test_gpu_ref.py -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -8 $O/pytest_gpu_$TAG.log
show4k() { python -c "import json;d=json.load(open('$1'));r=d['roofline'];print('$2', round(d['value'],1),'fps', round(d['ms_per_step'],3),'ms  lm', round(r['launch_ms'],3), 'frac', round(r['frac'],3), r['linearisations_per_level'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'err', d['pose_error_vs_gt_m'])"; }
timeout 300 python bench.py --workload 4k --steps 10 --no-cpu-baseline > $O/b_${TAG}_4k.json 2> $O/b_${TAG}_4k.err; show4k $O/b_${TAG}_4k.json "4k dense path"
ODO_LM_NO_DENSE=1 timeout 300 python bench.py --workload 4k --steps 10 --no-cpu-baseline --no-e2e > $O/b_${TAG}_4k_list.json 2> $O/b_${TAG}_4k_list.err; show4k $O/b_${TAG}_4k_list.json "4k list path"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -c 60 --csv --log-file $O/launches_${TAG}_4k.csv \
  python bench.py --workload 4k --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches_${TAG}_4k.log 2>&1
grep k_lm_solve $O/launches_${TAG}_4k.csv | tail -6 | cut -c 1-260
export ODO_BARRIER_TIMEOUT_MS=900000
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --log-file $O/sanitizer_${TOOL}_$TAG.txt \
    python -m pytest tests/test_gpu_groups.py tests/test_gpu_tracker.py -m gpu -x -q -k "group_sizes or batched_sequences or u8_ingest" > $O/sanitizer_${TOOL}_$TAG.pytest.log 2>&1
  echo "$TOOL rc=$?"; tail -3 $O/sanitizer_${TOOL}_$TAG.pytest.log; tail -4 $O/sanitizer_${TOOL}_$TAG.txt
done
