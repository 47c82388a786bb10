"""Prints the launches of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` launch list of a training loop
(everything after the second-to-last optimizer launch).  usage: python tools/launch_step.py launches.csv [marker=rmsprop]"""
import csv, sys
marker = sys.argv[2] if len(sys.argv) > 2 else "rmsprop"
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        v = float(x["Metric Value"]) / (1000.0 if x["Metric Unit"] in ("nsecond", "ns") else 1.0)
        rows.append((x["Kernel Name"].replace("ga3c::<unnamed>::", "").replace("ga3c::", "").replace("void ", "")[:64], x["Grid Size"], x["Block Size"], v))
idx = [i for i, r in enumerate(rows) if marker in r[0]]
start = idx[-2] + 1 if len(idx) > 1 else 0
tot = 0.0
print("| kernel | grid | block | us |\n|---|---|---|---|")
for r in rows[start:idx[-1] + 1]:
    print("| `%s` | %s | %s | %.1f |" % r)
    tot += r[3]
print(f"\n{idx[-1] + 1 - start} launches, {tot:.1f} us in all (each launch timed alone by ncu: serialised, cold caches)")
