#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_long.py tests/test_gpu_groups.py tests/test_gpu_camera.py tests/test_shard.py -m gpu -x -q -s 2>&1 | grep "GPU vs\|passed\|failed\|Error" | tail -3
python bench.py --seqs 1 --steps 50 --no-cpu-baseline --no-e2e > $O/t.json 2>/dev/null; python -c "import json;d=json.load(open('$O/t.json'));print('S1 pipelined', round(d['ms_per_step'],3), 'sync latency', d['latency_s1_ms'])"
python bench.py --seqs 8 --steps 50 --no-cpu-baseline --no-e2e --no-extras > $O/t.json 2>/dev/null; python -c "import json;d=json.load(open('$O/t.json'));print('S8', round(d['ms_per_step'],3))"
python bench.py --seqs 64 --steps 40 --no-cpu-baseline --no-extras > $O/t.json 2>/dev/null; python -c "import json;d=json.load(open('$O/t.json'));print('S64', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
