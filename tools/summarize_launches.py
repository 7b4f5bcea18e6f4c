"""ncu launch list (csv or csv.gz from tools/run_bench_and_profile.sh) -> markdown table + GEMM-class DRAM traffic.
   python tools/summarize_launches.py profiles/r01_launches_v2.csv.gz profiles/r01_launches_v2.md profiles/r01_gemm_traffic.json"""
import collections, csv, gzip, io, json, sys

src, out_md, out_json = sys.argv[1:4]
raw = gzip.open(src, "rt").read() if src.endswith(".gz") else open(src).read()
rows = [r for r in csv.reader(io.StringIO(raw)) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
per, names = collections.defaultdict(dict), {}
for r in rows[1:]:
    per[r[iid]][r[im]] = float(r[iv].replace(",", ""))
    names[r[iid]] = r[ik]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k, m in per.items():
    n = names[k].replace("void ", "").replace("ovmr::<unnamed>::", "").split("(")[0][:70]
    a = agg[n]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0) / 1e3
    a[2] += m.get("dram__bytes_read.sum", 0.0)
    a[3] += m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
lines = [f"# ncu launch list: {len(per)} launches of `bench.py --classes 32 --queries 1024 --steps 1 --warmup 3` "
         f"(timed step; `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`)",
         "", "Per-launch times are cold-cache and serialised under ncu: compare SHARES with bench.py's "
         "`roofline.kernel_ms_per_step`, not absolutes.", "",
         "| kernel | launches | total us | share | DRAM read MB/launch | DRAM write MB/launch |", "|---|---:|---:|---:|---:|---:|"]
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"| `{n}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[2] / a[0] / 1e6:.1f} | {a[3] / a[0] / 1e6:.1f} |")
cls = {"gemm": 0.0, "attention": 0.0, "layernorm": 0.0, "other": 0.0}
for n, a in agg.items():
    key = "gemm" if n.startswith("gemm_tn") else "attention" if n.startswith("attention") else \
        "layernorm" if n.startswith("layernorm") else "other"
    cls[key] += a[1]
lines += ["", "Class shares: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in cls.items())]
open(out_md, "w").write("\n".join(lines) + "\n")
g = [(a[0], a[2] + a[3]) for n, a in agg.items() if n.startswith("gemm_tn")]
n_l, bytes_ = sum(x[0] for x in g), sum(x[1] for x in g)
json.dump({"kernel_class": "gemm_tn_kernel / gemm_tn_pair_kernel (all GEMM launches of the timed step)",
           "launches": n_l, "dram_bytes_total": bytes_, "dram_bytes_per_launch": bytes_ / max(1, n_l),
           "source": src, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu launch list; "
                                  "ncu flushes caches between launches, so cross-kernel L2 reuse is not visible here"},
          open(out_json, "w"), indent=1)
print(open(out_md).read()[:1500])
