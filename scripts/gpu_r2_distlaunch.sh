#!/bin/bash
mkdir -p gpurun_out
timeout 80 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_dist_step.csv python scripts/dist_launches.py > gpurun_out/r2_launches_dist_step.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2_launches_dist_step.csv") if l.startswith('"'))]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    c = tot.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += v
s = sum(v for _, v in tot.values())
print(f"{sum(c for c, _ in tot.values())} launches, {s / 1e6:.3f} ms (ncu per-launch times: cold, serialised)")
for k, (c, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{k:40s} {c:3d} launches {v / 1e3:10.1f} us  {100 * v / s:5.1f} %")
PY
