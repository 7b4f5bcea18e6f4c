"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, re, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("ovmr::<unnamed>::", "")
    name = re.sub(r"at::native::.*?(\w+_kernel\w*).*", r"torch:\1", name)
    tot[name][0] += 1
    tot[name][1] += float(r[14]) / 1e3
total = sum(v[1] for v in tot.values())
print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:90]}` | {n} | {us:.1f} | {100*us/total:.1f}% |")
print(f"| **all** | {len(rows)} | {total:.1f} | 100% |")
