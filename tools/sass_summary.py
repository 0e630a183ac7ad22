"""Counts of the Blackwell-specific SASS mnemonics per kernel of libols_b200.so (no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st (TMEM), UTMALDG / UTMASTG = TMA load / store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, FFMA2 = packed fp32x2 FMA, MUFU.EX2 = ex2.approx, REDUX = redux.sync."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "online_lang_splatting_b200", "lib", "libols_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "FFMA2", "FMUL2", "MUFU.EX2", "REDUX", "LDGSTS", "ATOMS", "RED.E"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    total[cur] += 1
    for p in pats:
        if re.search(r"\b" + re.escape(p), line):
            counts[cur][p] += 1
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n
print("# cuobjdump -sass online_lang_splatting_b200/lib/libols_b200.so (sm_100a), instruction counts per kernel")
print("# " + " ".join(pats))
for k, c in counts.items():
    if not any(c.values()):
        continue
    name = demangle(k)
    name = re.sub(r"\(.*", "", name)[:72]
    print(f"{name:74s} sass={total[k]:6d}  " + " ".join(f"{p}={c[p]}" for p in pats if c[p]))
