#!/bin/bash
# Runs scripts/tc_check.py for every tensor-core kernel version in its own process (a trap in one
# version must not poison the others), prints a table, and echoes the fastest PASSing version.
best=1; best_rate=0
for v in ${TC_VERSIONS:-1 2 3 4}; do
  NANN_TC_KERNEL=$v timeout 300 python scripts/tc_check.py > gpurun_out/tc_check_v$v.log 2>&1
  rc=$?
  worst=$(grep WORST gpurun_out/tc_check_v$v.log | awk '{print $2, $3}')
  rate=$(grep "^tensor:" gpurun_out/tc_check_v$v.log | awk '{print $2}')
  echo "tc kernel v$v: rc=$rc worst=[$worst] tensor=${rate:-0} M rows/s" >&2
  if grep -q "PASS" gpurun_out/tc_check_v$v.log && [ -n "$rate" ]; then
    if python -c "import sys; sys.exit(0 if float('$rate') > float('$best_rate') else 1)"; then best=$v; best_rate=$rate; fi
  else
    tail -5 gpurun_out/tc_check_v$v.log >&2
  fi
done
echo $best
