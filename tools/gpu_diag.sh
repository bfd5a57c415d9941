#!/bin/bash
# First-contact diagnostics: every test FUNCTION of the -m gpu files in its own process (no -x), so one
# device trap does not hide the rest.  Logs and a summary land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag_gpu.txt 2>&1
summary=gpurun_out/diag_summary.txt
: > $summary
for f in "$@"; do
  funcs=$(grep -oE "^def (test_[a-zA-Z0-9_]+)" "$f" | awk '{print $2}')
  for fn in $funcs; do
    log="gpurun_out/diag_$(basename $f .py)_${fn}.log"
    timeout 600 python -m pytest "$f" -q -m gpu -k "$fn" --timeout 240 -p no:cacheprovider > "$log" 2>&1
    echo "$fn rc=$? $(tail -n 1 $log)" | tee -a $summary
  done
done
grep -h -E "^(FAILED|ERROR)|assert|Error|error:|mbarrier timeout" gpurun_out/diag_*.log | head -80 >> $summary
tail -n 60 $summary
