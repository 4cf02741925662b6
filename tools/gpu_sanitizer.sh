#!/bin/bash
# compute-sanitizer passes over the persistent group kernels (k_refine, k_lm_solve) and the tracker's three-stream frame
# pipeline.  gpurun --timeout 2400 -- 'bash tools/gpu_sanitizer.sh r2x'
# The barrier time-out is raised because the sanitizer slows the kernels by 100x and more; the selected tests are the
# ones that drive every group size (1, 2, cluster and counter-barrier groups) and batched sequences with u8 ingest.
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
export ODO_BARRIER_TIMEOUT_MS=900000
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --log-file $O/sanitizer_${TOOL}_$TAG.txt \
    python -m pytest tests/test_gpu_groups.py tests/test_gpu_tracker.py -m gpu -x -q \
      -k "group_sizes or batched_sequences or u8_ingest" > $O/sanitizer_${TOOL}_$TAG.pytest.log 2>&1
  echo "$TOOL rc=$?"; tail -3 $O/sanitizer_${TOOL}_$TAG.pytest.log; tail -4 $O/sanitizer_${TOOL}_$TAG.txt
done
