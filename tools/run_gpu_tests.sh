#!/bin/bash
# Run every -m gpu test file in its own process (a device trap in one file must not poison the next)
# and keep the logs under gpurun_out/.  Usage: tools/run_gpu_tests.sh [pytest args]
mkdir -p gpurun_out
rc_all=0
for f in tests/test_gpu_*.py tests/test_octree_conv.py; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -q -m gpu -x --timeout 300 "$@" > "gpurun_out/${name}.log" 2>&1
  rc=$?
  echo "== $name rc=$rc"; tail -n 25 "gpurun_out/${name}.log"
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
