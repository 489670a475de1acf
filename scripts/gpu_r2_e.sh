#!/bin/bash
# round-2 call E (1 GPU): full gpu test-suite, smoke, builder at 10M, gather roofline (events + ncu), racecheck of the tc8 scorer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python scripts/builder_bench.py 10000000 1 2>&1 | tail -2
timeout 300 python scripts/gather_bench.py 2>&1 | tail -4
NCU_ONE=1 timeout 600 ncu --set full --clock-control none -k regex:gather_rows_vec_kernel -c 2 -o gpurun_out/r2_ncu_gather -f python scripts/gather_bench.py > gpurun_out/r2_ncu_gather.log 2>&1; echo "ncu gather rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_ncu_gather_summary.json gpurun_out/r2_ncu_gather.ncu-rep > /dev/null 2>&1; echo "summary rc=$?"
# racecheck with the mbarrier watchdog compiled out (the instrumentation slows the tcgen05 kernels ~100x)
cp nann_b200/lib/libnann_b200.so /tmp/libnann_b200.so.keep
NANN_NVCC_EXTRA="-DNANN_MBAR_WATCHDOG_NS=0" python nann_b200/build.py --force > /dev/null 2>&1; echo "rebuild rc=$?"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -c "Race reported\|hazard" gpurun_out/r2_racecheck.log; grep "RACECHECK SUMMARY\|tensor ok\|^ok" gpurun_out/r2_racecheck.log
grep "Race reported" gpurun_out/r2_racecheck.log | sed 's/0x[0-9a-f]*//g' | cut -c1-220 | sort | uniq -c | sort -rn | head -12
cp /tmp/libnann_b200.so.keep nann_b200/lib/libnann_b200.so
