"""Summarise an .ncu-rep: key raw metrics per kernel and the hottest SASS lines (stall samples).
usage: python tools/ncu_summary.py file.ncu-rep [n_hot]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_red.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("==== %s  grid %s block %s" % (r[hdr.index("Kernel Name")][:60], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print("  %-88s %s %s" % (k, r[i], units[i]))
    st = sorted(((float(r[hdr.index(h)] or 0), h.split("stalled_")[1].split("_per_")[0]) for h in stall), reverse=True)[:7]
    print("  stalls/issue: " + ", ".join("%s %.2f" % (n, v) for v, n in st))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1] if i > 0 and len(rows[i - 1]) > 1 else "?" for i in secs]
for si, h in enumerate(secs):
    hd = rows[h]; isrc = hd.index("Source"); ins = hd.index("# Samples")
    sc = [i for i, c in enumerate(hd) if c.startswith("stall_") and "Not Issued" not in c]
    end = secs[si + 1] - 1 if si + 1 < len(secs) else len(rows)
    items, tot = [], 0
    for r in rows[h + 1:end]:
        try: s = int(r[ins] or 0)
        except Exception: continue
        tot += s
        items.append((s, r[isrc][:58], sorted(((int(r[i]), hd[i][6:]) for i in sc if r[i] and int(r[i]) > 0), reverse=True)[:3]))
    print("---- hot SASS: %s (samples %d)" % (names[si][:70], tot))
    for s, t, st in sorted(items, key=lambda x: -x[0])[:nhot]:
        print("%6d %-58s %s" % (s, t, st))
