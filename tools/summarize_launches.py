"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean
duration and share of the total of OUR kernels (names starting with ga3c::).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/xxx.md"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append((r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"])))
agg = collections.OrderedDict()
for name, grid, block, ns in rows:
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    short = re.sub(r"ga3c::(anonymous namespace)::", "ga3c::", short)
    key = (short, grid, block)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += ns
ours = sum(v[1] for k, v in agg.items() if "ga3c" in k[0])
print(f"| kernel | grid | block | launches | mean us | share of ga3c kernel time |")
print("|---|---|---|---|---|---|")
for (short, grid, block), (n, tot) in agg.items():
    share = f"{tot/ours:.3f}" if "ga3c" in short else "-"
    print(f"| `{short}` | {grid} | {block} | {n} | {tot/n/1e3:.2f} | {share} |")
