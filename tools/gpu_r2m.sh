#!/bin/bash
O=gpurun_out
for G in 29 16 8; do
  echo "== S=1 G=$G"; ODO_LM_CTAS_PER_SEQ=$G python tools/lm_cycles.py 1 2>&1 | grep "^frame" | tail -1
done
echo "== S=1 G=16 no cluster"; ODO_GROUP_NO_CLUSTER=1 ODO_LM_CTAS_PER_SEQ=16 python tools/lm_cycles.py 1 2>&1 | grep "^frame" | tail -1
for G in 9 8; do echo "== S=8 G=$G"; ODO_LM_CTAS_PER_SEQ=$G python tools/lm_cycles.py 8 2>&1 | grep "^frame" | tail -1; done
echo "== S=8 G=8 no cluster"; ODO_GROUP_NO_CLUSTER=1 ODO_LM_CTAS_PER_SEQ=8 python tools/lm_cycles.py 8 2>&1 | grep "^frame" | tail -1
ODO_LM_CTAS_PER_SEQ=8 timeout 600 python -m pytest tests/test_gpu_lm.py tests/test_gpu_tracker.py tests/test_gpu_groups.py tests/test_gpu_long.py -m gpu -x -q -s 2>&1 | grep "GPU vs\|passed\|failed\|Error" | tail -4
B="python bench.py --steps 40 --no-cpu-baseline --no-e2e --no-extras"
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
for S in 1 8 16 32; do $B --seqs $S > $O/t.json 2>/dev/null; show $O/t.json "S$S default"; done
ODO_LM_CTAS_PER_SEQ=16 $B --seqs 1 > $O/t.json 2>/dev/null; show $O/t.json "S1 G=16 cluster"
ODO_LM_CTAS_PER_SEQ=8 $B --seqs 1 > $O/t.json 2>/dev/null; show $O/t.json "S1 G=8 cluster"
ODO_LM_CTAS_PER_SEQ=8 $B --seqs 8 > $O/t.json 2>/dev/null; show $O/t.json "S8 G=8 cluster"
