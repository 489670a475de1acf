#!/bin/bash
# round-2 call G (1 GPU): builder after the survivor-queue change (tests + 1M/10M timing), configs[4] sweep with CUDA graphs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_builder_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/builder_bench.py 1000000 2 2>&1 | tail -2
timeout 900 python scripts/builder_bench.py 10000000 1 2>&1 | tail -1
DUR=2 timeout 900 python scripts/sweep.py 2>&1 | grep -v "^\[bench\]" | tail -10
cp gpurun_out/sweep.json gpurun_out/r2_sweep.json
