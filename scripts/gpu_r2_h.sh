#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_builder_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/builder_bench.py 1000000 2 2>&1 | tail -2
timeout 900 python scripts/builder_bench.py 10000000 1 2>&1 | tail -1
