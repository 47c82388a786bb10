#!/bin/bash
# quick GPU iteration: per-layer parity at a few batch sizes, then the bench (no CPU leg), short summary
timeout 300 python tests/gpu_layer_report.py 3 33 300 > gpurun_out/debug_iter.log 2>&1; echo "debug rc=$?"
grep -E "^---|  p  |  v  |n1 |n2 |dn1|dn2|grad/conv11/w|grad/conv12/w|grad/dense1/w" gpurun_out/debug_iter.log | awk '{printf "%s ", $0; if (NR%10==0) print ""}'; echo
timeout 400 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_iter.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_iter.json"))
print("TPS %.3fM ms %.4f | PPS %.3fM ms %.4f | e2e %.0f / %.0f" % (d["value"]/1e6,d["ms_per_step"],d["pps"]["value"]/1e6,d["pps"]["ms_per_step"],d["e2e"]["value"],d["pps"]["e2e"]["value"]))
for k,v in d["roofline"]["kernels_train"].items(): print("train %-14s %8.2f us  frac %.3f %s" % (k,v["avg_us"],v["frac"],v["unit"]))
for k,v in d["roofline"]["kernels_predict"].items(): print("pred  %-14s %8.2f us  frac %.3f %s" % (k,v["avg_us"],v["frac"],v["unit"]))
print(d["clocks"])
PY
