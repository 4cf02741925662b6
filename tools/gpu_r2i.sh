#!/bin/bash
O=gpurun_out
B="python bench.py --steps 40 --no-cpu-baseline --no-e2e --no-extras"
show() { python -c "import json,sys;d=json.load(open('$1'));print('$2', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"; }
for P in 1 0; do for G in 1 2; do
  ODO_LM_STREAM_PRIORITY=$P ODO_LM_CTAS_PER_SEQ=$G $B --seqs 64 > $O/t.json 2>/dev/null; show $O/t.json "S64 prio=$P G=$G"
done; done
for P in 1 0; do
  ODO_LM_STREAM_PRIORITY=$P $B --seqs 32 > $O/t.json 2>/dev/null; show $O/t.json "S32 prio=$P"
  ODO_LM_STREAM_PRIORITY=$P $B --seqs 8 > $O/t.json 2>/dev/null; show $O/t.json "S8 prio=$P"
done
