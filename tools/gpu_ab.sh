#!/bin/bash
# A/B visit: parity tests, then the HBM-resident bench leg with an environment switch off and on.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_ab.sh TAG ENVVAR'
TAG=${1:-ab}
VAR=${2:-ODO_DISP_NOPRUNE}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
tail -4 $O/pytest_gpu_$TAG.log
for v in 0 1; do
  if [ $v = 1 ]; then export $VAR=1; fi
  timeout 300 python bench.py --no-cpu-baseline --no-e2e > $O/bench_${TAG}_$v.json 2> $O/bench_${TAG}_$v.err
  python - <<PY
import json
d = json.load(open("$O/bench_${TAG}_$v.json"))
print("$VAR=$v", round(d["value"]), "frames/s", d["ms_per_step"], {k: round(x, 3) for k, x in d["roofline"]["stage_ms_per_step"].items()})
PY
done
