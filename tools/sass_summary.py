"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (cuobjdump -sass of the built library):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store), UTCBAR (tcgen05.commit),
HMMA (legacy mma.sync), MUFU.EX2.   python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ovmr_b200", "libovmr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "MUFU.EX2", "USETMAXREG"]
rows, cur, counts = [], None, None
names = iter(demangle)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur is not None:
            rows.append((cur, counts))
        cur, counts = next(names), collections.Counter()
        continue
    if cur is None:
        continue
    for k in keys:
        if re.search(r"\b" + re.escape(k), line):
            counts[k] += 1
if cur is not None:
    rows.append((cur, counts))


def short(n):
    n = re.sub(r"ovmr::\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(CUtensorMap_st.*", "", n)
    n = re.sub(r"^void ", "", n)
    return n[:70]


print("# r02 — SASS evidence per kernel (`cuobjdump -sass ovmr_b200/libovmr_b200.so`, sm_100a; `tools/sass_summary.py`)\n")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor load / store,")
print("UTMAPF = TMA L2 prefetch, HMMA = legacy mma.sync, USETMAXREG = setmaxnreg.  Kernels without any of these (row kernels, head,")
print("preprocessing, backward element-wise kernels) are omitted.\n")
print("| kernel | " + " | ".join(keys) + " |")
print("|---|" + "---:|" * len(keys))
for name, c in sorted(rows, key=lambda r: short(r[0])):
    if not any(c[k] for k in keys if k != "MUFU.EX2"):
        continue
    print(f"| `{short(name)}` | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print("\nTotals: " + ", ".join(f"{k} {tot[k]}" for k in keys))
