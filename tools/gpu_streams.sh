for t in 1 2 4; do timeout 300 python bench.py --streams $t --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams $t', round(d['value']), d['ms_per_step'])"; done
for s in 128 256; do timeout 300 python bench.py --seqs $s --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('seqs $s', round(d['value']), d['ms_per_step'], {k: round(x,3) for k,x in d['roofline']['stage_ms_per_step'].items()})"; done
