#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_pyramid.py tests/test_gpu_depth.py tests/test_gpu_tracker.py tests/test_gpu_long.py tests/test_gpu_groups.py tests/test_gpu_ref.py -m gpu -x -q -s 2>&1 | grep "GPU vs\|passed\|failed\|Error" | tail -4
B="python bench.py --steps 40 --no-cpu-baseline --no-e2e --no-extras"
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
$B --seqs 64 > $O/t.json 2>/dev/null; show $O/t.json "S64 staged blur"
ODO_BLUR_GATHER=1 $B --seqs 64 > $O/t.json 2>/dev/null; show $O/t.json "S64 gather blur"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_blur3" -c 12 --csv --log-file $O/launches_blur.csv \
  python bench.py --seqs 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
grep "gpu__time_duration" $O/launches_blur.csv | awk -F'","' '{print $5, $NF}' | tail -4
