"""The C-ABI library: loads, exports every symbol include/nann_b200.h declares, refuses to compute
without a GPU (no CPU fallback), and its host-side npy/HugeConst loader matches numpy + the oracle.
CPU only -- no kernels are launched here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nann_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(nann_[a-z0-9_]+)\s*\(", src))
    for stem in re.findall(r"\b(nann_[a-z0-9_]+_)##SFX\s*\(", src):      # NANN_RAGGED_DECL(T, SFX) expansions
        names |= {stem + "i32", stem + "i64"}
    names = {n for n in names if not n.endswith("_")}
    names -= {"nann_alloc_fn"}
    return sorted(names)


def test_build_and_exports():
    from nann_b200 import build
    so = build.build()
    assert os.path.exists(so)
    lib = C.CDLL(so)
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.nann_abi_version() == 2


def test_library_targets_sm_100a():
    from nann_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use oracle/."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "nann_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".cc")):
                s = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"(import|from)\s+oracle|nann_oracle|orc_[a-z]+\(|libnann_oracle", s):
                    bad.append(f)
    assert not bad, bad


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_gpu():
    import nann_b200
    with pytest.raises(nann_b200.NannError) as e:
        nann_b200.top_k(np.arange(8, dtype=np.float32), 2)
    assert e.value.code == nann_b200._lib.FAILED_PRECONDITION
    with pytest.raises(nann_b200.NannError):
        nann_b200.group_gather(np.arange(3, dtype=np.int64), [0, 3], [0], [0, 1])
    with pytest.raises(nann_b200.NannError):
        nann_b200.Scorer.mlp(np.zeros((512, 256)), np.zeros(512), np.zeros((512, 512)), np.zeros(512), np.zeros(512))


def test_missing_library_is_an_import_error(tmp_path, monkeypatch):
    from nann_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.lib()


def test_npy_loader_host_side(tmp_path, oracle):
    """HugeConst ctor semantics (huge_const_op.cc:85-182), host-only handle (device=-1)."""
    import nann_b200
    from nann_b200 import _lib
    cases = [np.array([[1, 2], [3, 4], [5, 6]], np.int32), np.array([[1, 2, 3], [4, 5, 6]], np.int64),
             np.array([[1, 2, 3, 4, 5, 6]], np.float32), np.random.default_rng(0).random((7, 5)).astype(np.float16),
             np.random.default_rng(1).random((3, 2, 2))]
    for n, a in enumerate(cases):
        p = str(tmp_path / f"h{n}.npy")
        np.save(p, a)
        assert nann_b200.ops.npy_peek(p) == (a.dtype, a.shape)
        h = nann_b200.huge_const(p, a.dtype, a.shape, device=-1)
        np.testing.assert_array_equal(h.numpy(), a)
        assert h.nbytes == a.nbytes and h.device_ptr is None
        st, _ = oracle.huge_const_check(p, a.dtype, a.shape)
        assert st == 0
        wrong = np.float64 if a.dtype != np.float64 else np.float32
        with pytest.raises(nann_b200.NannError) as e:
            nann_b200.huge_const(p, wrong, a.shape, device=-1)
        assert e.value.code == _lib.INTERNAL == oracle.huge_const_check(p, wrong, a.shape)[0]
        with pytest.raises(nann_b200.NannError) as e:
            nann_b200.huge_const(p, a.dtype, (a.shape[0] + 1,) + a.shape[1:], device=-1)
        assert e.value.code == _lib.INTERNAL
    with pytest.raises(nann_b200.NannError) as e:
        nann_b200.huge_const(str(tmp_path / "missing.npy"), np.float32, (1,), device=-1)
    assert e.value.code == _lib.NOT_FOUND
    np.save(str(tmp_path / "f.npy"), np.asfortranarray(np.ones((2, 3), np.float32)))
    with pytest.raises(nann_b200.NannError) as e:
        nann_b200.huge_const(str(tmp_path / "f.npy"), np.float32, (2, 3), device=-1)
    assert e.value.code == _lib.UNIMPLEMENTED
    # npy format 2.0 header
    a = np.arange(12, dtype=np.int64).reshape(3, 4)
    with open(str(tmp_path / "v2.npy"), "wb") as f:
        np.lib.format.write_array(f, a, version=(2, 0))
    np.testing.assert_array_equal(nann_b200.huge_const(str(tmp_path / "v2.npy"), np.int64, (3, 4), device=-1).numpy(), a)


def test_header_is_plain_c_and_links(tmp_path):
    """include/nann_b200.h compiles as strict C99 (no C++ or torch types in the ABI) and a C program links against
    the library; without a GPU a compute entry point answers FAILED_PRECONDITION instead of falling back."""
    from nann_b200 import build
    so = build.build()
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include "nann_b200.h"
int main(void) {
  float in[4] = {1.f, 3.f, 2.f, 0.f}, val[2];
  int32_t idx[2];
  nann_status st;
  if (nann_abi_version() != NANN_B200_ABI_VERSION) return 2;
  st = nann_topk_v2_f32(in, 1, 4, 2, 1, val, idx, NULL);
  printf("%d %s\n", (int)st, st == NANN_OK ? "ok" : nann_last_error());
  if (st == NANN_OK) return (idx[0] == 1 && idx[1] == 2) ? 0 : 3;
  return st == NANN_FAILED_PRECONDITION ? 0 : 4;
}
''')
    exe = tmp_path / "abi"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), so, "-Wl,-rpath," + os.path.dirname(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
