#!/bin/bash
# round-2 call D (1 GPU): new op tests, nvlink counter format, builder launch list + 10M timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_scorer_gpu.py -m gpu -x -q -k "bloom or admission" 2>&1 | tail -8
nvidia-smi nvlink -gt d -i 0 2>&1 | head -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_builder_launches.csv python scripts/builder_bench.py 1000000 1 > gpurun_out/r2_builder_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_builder_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:60], []).append(float(r[vi].replace(",", "")))
    except Exception: pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} n={len(v):4d} sum={sum(v)/1e6:9.2f} ms  {100*sum(v)/tot:5.1f}%")
PY
timeout 900 python scripts/builder_bench.py 10000000 1 2>&1 | tail -3
