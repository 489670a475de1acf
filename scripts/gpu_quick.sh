#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== timeline v3"; NANN_TC_KERNEL=3 timeout 300 python scripts/tc_timeline.py 2>&1 | tail -80 | tee gpurun_out/tc_timeline_v3.log
echo "== pytest gpu"; NANN_TC_KERNEL=3 timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
