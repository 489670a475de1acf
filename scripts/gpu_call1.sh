#!/bin/bash
set -u
mkdir -p gpurun_out
TC_VERSIONS="3 7" bash scripts/gpu_quick.sh 2>&1 | tail -60
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
