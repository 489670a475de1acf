"""Summarise .ncu-rep captures (ncu --set full) into a small JSON for profiles/.
usage: python scripts/ncu_summary.py OUT.json REP [REP...]"""
import csv, io, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct"]

def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)

out = {}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    ks = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT: d[h] = f"{r[i]} {units[i]}".strip()
        try:
            rd = hdr.index("dram__bytes_read.sum"); wr = hdr.index("dram__bytes_write.sum")
            d["dram_bytes_total"] = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
        except ValueError:
            pass
        ks.append(d)
    out[rep.split("/")[-1].replace(".ncu-rep", "")] = ks
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
