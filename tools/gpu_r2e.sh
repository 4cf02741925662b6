#!/bin/bash
TAG=${1:-r2e}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -6 $O/pytest_gpu_$TAG.log
show4k() { python -c "import json;d=json.load(open('$1'));r=d['roofline'];print('$2', round(d['value'],1),'fps', round(d['ms_per_step'],3),'ms  lm', round(r['launch_ms'],3), 'frac', round(r['frac'],3), r['linearisations_per_level'], 'err', d['pose_error_vs_gt_m'])"; }
timeout 300 python bench.py --workload 4k --steps 10 --no-cpu-baseline --no-e2e > $O/b_${TAG}_4k.json 2> $O/b_${TAG}_4k.err; show4k $O/b_${TAG}_4k.json "4k dense path"
B="python bench.py --steps 40 --no-cpu-baseline --no-e2e --no-extras"
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
$B --seqs 64 > $O/b_${TAG}_64.json 2>/dev/null; show $O/b_${TAG}_64.json "S64"
$B --seqs 16 > $O/b_${TAG}_16.json 2>/dev/null; show $O/b_${TAG}_16.json "S16"
export ODO_BARRIER_TIMEOUT_MS=900000
timeout 600 compute-sanitizer --tool racecheck --log-file $O/sanitizer_racecheck_lm_$TAG.txt \
    python -m pytest tests/test_gpu_groups.py -m gpu -x -q -k "lm_group_sizes" > $O/sanitizer_racecheck_lm_$TAG.pytest.log 2>&1
echo "racecheck rc=$?"; tail -2 $O/sanitizer_racecheck_lm_$TAG.pytest.log; tail -3 $O/sanitizer_racecheck_lm_$TAG.txt
