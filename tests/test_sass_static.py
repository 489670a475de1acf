"""Static checks on the built library (no GPU): the fatbin holds sm_100a code only, the scorer and the index builder's
k-NN kernel really are tcgen05 / TMEM / TMA kernels (SASS mnemonics of /opt/skills/guides/B200_PROFILING.md: UTCHMMA =
tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = bulk copy through the TMA engine, UTCBAR = tcgen05.commit, SYNCS = mbarrier), no
kernel spills to local memory except the two that index small per-thread arrays, and the exchange kernels use
system-scope fences."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "nann_b200", "lib", "libnann_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def sass():
    import __graft_entry__ as g
    g.build()
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    per = collections.defaultdict(collections.Counter)
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
            body = line.split("*/", 1)[1]
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", body)
            if m:
                per[cur][m.group(1)] += 1
                per[cur][m.group(1) + m.group(2)] += 1
    return out, per


def _kernel(per, needle):
    hits = [k for k in per if needle in k]
    assert hits, f"no kernel matching {needle}"
    return hits


def test_fatbin_is_sm_100a_only(sass):
    out, _ = sass
    assert set(re.findall(r"arch = (sm_\w+)", out)) == {"sm_100a"}


def test_scorer_is_a_tcgen05_tmem_tma_kernel(sass):
    _, per = sass
    for k in _kernel(per, "mlp_tc8_kernel"):
        c = per[k]
        assert c["UTCHMMA"] >= 8, c          # tcgen05.mma kind::f16 (hi and lo stages of the two layers)
        assert c["LDTM"] >= 8                # tcgen05.ld: accumulators out of TMEM in both epilogues
        assert c["UBLKCP"] >= 4              # weight image + DSMEM hand-off through the bulk-copy engine
        assert c["UTCBAR"] >= 4              # tcgen05.commit onto mbarriers (incl. the multicast form for the pair)
        assert c["SYNCS"] >= 20              # mbarrier pipeline
        assert c["HMMA"] == 0                # no legacy mma.sync path


def test_builder_knn_is_a_tcgen05_kernel(sass):
    _, per = sass
    for k in _kernel(per, "knn_filter_kernel"):
        c = per[k]
        assert c["UTCHMMA"] >= 4 and c["LDTM"] >= 1 and c["UBLKCP"] >= 1 and c["UTCBAR"] >= 1


def test_exchange_kernels_use_system_scope_ordering(sass):
    """stores into peer windows are followed by a system-scope fence / release store before the flag"""
    _, per = sass
    for name in ("dist_bucket_push_kernel", "dist_return_kernel", "dist_bcast_kernel", "shard_merge_kernel", "topk_kernel"):
        for k in _kernel(per, name):
            assert per[k]["MEMBAR.ALL.SYS"] + per[k]["MEMBAR.SC.SYS"] >= 1, (name, per[k])


def test_no_unexpected_local_memory(sass):
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True, check=True).stdout
    stack = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and cur:
            stack[cur] = (int(m.group(1)), int(m.group(2)))
            cur = None
    assert len(stack) >= 70
    spilling = {k: v for k, v in stack.items() if v[1] > 0}
    allowed = ("knn_filter_kernel", "knn_refine_kernel", "bloom_claim_kernel")     # small per-thread arrays, by design
    assert all(any(a in k for a in allowed) for k in spilling), spilling
    for k, (regs, _) in stack.items():
        if "mlp_tc8_kernel" in k:
            assert regs <= 168                 # 320 threads x 168 registers is the budget of one CTA per SM
