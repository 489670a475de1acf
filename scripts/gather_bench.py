"""HBM-gather roofline run (BASELINE configs[2] table): nann_gather_rows (GatherV2, build_opt_graph.py:92) over random
512-B rows of a 10M x 128 f32 table (5.1 GB, far beyond the 126 MB L2).  CUDA-event timing after warm-up; achieved =
algorithmic bytes (512 B read + 512 B written per row) / time, against the measured copy bandwidth of MEASURED_PEAKS.json.
With NCU_ONE=1 only one warm call + one measured call run (for `ncu --set full -k regex:gather_rows_vec_kernel`)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nann_b200  # noqa: F401  (loads the library)
from nann_b200 import _lib

n_rows, d = int(os.environ.get("ROWS", 10_000_000)), 128
one = os.environ.get("NCU_ONE") == "1"
table = torch.randn(n_rows, d, device="cuda")
pk = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
res = {}
L = _lib.lib()
for n in ((1 << 24,) if one else (1 << 20, 1 << 22, 1 << 24)):
    ids = torch.randint(0, n_rows, (n,), device="cuda", dtype=torch.int32)
    out = torch.empty(n, d, device="cuda")
    args = (C.c_void_p(table.data_ptr()), n_rows, d * 4, C.c_void_p(ids.data_ptr()), n, C.c_void_p(out.data_ptr()), None)
    for _ in range(1 if one else 5):
        _lib.check(L.nann_gather_rows(*args))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 1 if one else 20
    e0.record()
    for _ in range(reps):
        _lib.check(L.nann_gather_rows(*args))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = n * d * 4 * 2 / ms / 1e6
    assert torch.equal(out[:1000], table[ids[:1000].long()])
    res[n] = {"ms": ms, "GBps_read_plus_write": gbs, "frac_of_measured_copy_peak": gbs / pk["hbm_gbs"],
              "algorithmic_bytes": n * d * 4 * 2}
    print(f"gather {n:9d} rows x 512 B: {ms:8.3f} ms  {gbs:8.1f} GB/s (r+w)  = {gbs / pk['hbm_gbs']:.3f} of measured copy peak", flush=True)
if not one:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"table_rows": n_rows, "row_bytes": 512, "peak_GBps": pk["hbm_gbs"], "results": res},
              open("gpurun_out/r2_gather_bench.json", "w"), indent=1)
