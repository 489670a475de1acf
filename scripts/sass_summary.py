"""Static evidence from the built library (no GPU needed): per-kernel registers / stack, and the Blackwell SASS
mnemonics (tcgen05 = UTC*, TMEM loads = LDTM, TMA-engine bulk copies = UBLKCP, mbarrier = SYNCS) per kernel.
usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nann_b200", "lib", "libnann_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
        cur = None
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTCCP|UTCATOMSWS|LDTM|STTM|UBLKCP|UBLKRED|UTMALDG|UTMASTG|SYNCS|HMMA|FFMA|MEMBAR|ATOM|RED|CCTL|ACQBULK|NANOSLEEP|LDG|STG|LDS|STS|REDUX|SHFL|BAR)\b")
counts = collections.defaultdict(collections.Counter)
n_ins = collections.Counter()
arch = set(re.findall(r"arch = (sm_\w+)", sass))
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
        n_ins[cur] += 1
        for op in pat.findall(line.split("*/", 1)[1]):
            counts[cur][op] += 1
print(f"library: nann_b200/lib/libnann_b200.so   arch in the fatbin: {sorted(arch)}   kernels: {len(usage)}")
print(f"{'kernel':44s} {'regs':>4s} {'stack':>5s} {'sass':>6s}  mnemonics")
for k in sorted(usage, key=lambda k: -n_ins[k]):
    r, st, _ = usage[k]
    ops = " ".join(f"{o}:{c}" for o, c in sorted(counts[k].items(), key=lambda x: -x[1]))
    print(f"{dem(k)[:44]:44s} {r:4d} {st:5d} {n_ins[k]:6d}  {ops}")
