#!/bin/bash
# lean 8-GPU visit: device-timed weak scaling and config 4 as written (strong), no CV_32F / reference legs
TAG=${1:-r2u}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 200 $TR bench.py --gpus 8 --seqs-total 64 --steps 40 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > $O/scale8_strong_$TAG.json 2> $O/scale8_strong_$TAG.err
timeout 200 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > $O/scale8_weak_$TAG.json 2> $O/scale8_weak_$TAG.err
python - <<PY
import json
for n in ('strong','weak'):
    for l in open('$O/scale8_%s_$TAG.json' % n):
        if l.startswith('{'):
            d=json.loads(l); print(n, 'N=8 value', round(d['value']), 'ms', round(d['ms_per_step'],3), d['scaling'], d['config']['sequences_per_gpu'])
PY
