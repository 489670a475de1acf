#!/bin/bash
set -u
mkdir -p gpurun_out
export TC_VERSIONS="${TC_VERSIONS:-8}"
bash scripts/gpu_quick.sh 2>&1 | tail -50
