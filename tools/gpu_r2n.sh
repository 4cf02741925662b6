#!/bin/bash
cp odometry_b200/libodo_b200.so /tmp/lib_base.so
echo "== base"; python tools/lm_cycles.py 1 2>&1 | grep "^frame" | tail -1
for V in "-DODO_QR_REDUX_PIVOT" "-DODO_QR_RCP_BACKSUB" "-DODO_QR_REDUX_PIVOT -DODO_QR_RCP_BACKSUB"; do
  ODO_EXTRA_NVCC="$V" python odometry_b200/build.py --force > /dev/null 2>&1
  echo "== $V"; python tools/lm_cycles.py 1 2>&1 | grep "^frame" | tail -1
  timeout 300 python -m pytest tests/test_gpu_long.py -m gpu -x -q -s -k oracle_wide 2>&1 | grep "GPU vs\|passed\|failed" | tail -2
done
cp /tmp/lib_base.so odometry_b200/libodo_b200.so
