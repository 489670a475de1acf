#!/bin/bash
mkdir -p gpurun_out
cp nann_b200/lib/libnann_b200.so /tmp/libnann_b200.so.keep
NANN_NVCC_EXTRA="-DNANN_MBAR_WATCHDOG_NS=0" python nann_b200/build.py --force > /dev/null 2>&1; echo "rebuild rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_r2.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/r2_memcheck.log
timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 9 python scripts/sanitize_r2.py > gpurun_out/r2_initcheck.log 2>&1; echo "initcheck rc=$?"
tail -4 gpurun_out/r2_initcheck.log
cp /tmp/libnann_b200.so.keep nann_b200/lib/libnann_b200.so
