"""Per-source-line stall samples from an .ncu-rep captured with --import-source on (-lineinfo build).
usage: python tools/ncu_lines.py file.ncu-rep [kernel substring] [n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath = func = None; hdr = None
agg = collections.defaultdict(lambda: [0, collections.Counter(), ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or want not in (func or ""): continue
    try: line = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr, r))
    try: s = int(d.get("# Samples") or 0)
    except ValueError: s = 0
    key = (func.split("(")[0], fpath.split("/")[-1], line)
    a = agg[key]; a[0] += s; a[2] = r[1][:110]
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v and v != "0":
            try: a[1][k[6:]] += int(v)
            except ValueError: pass
tot = collections.Counter()
for (f, _, _), a in agg.items(): tot[f] += a[0]
for f in tot:
    print(f"==== {f}: {tot[f]} samples")
    for (ff, fp, line), a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if ff != f or a[0] == 0: continue
        n -= 1
        if n < 0: break
        print(f"{a[0]:6d} {fp}:{line:<4d} {dict(a[1].most_common(2))}  | {a[2].strip()}")
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
