#!/bin/bash
# GPU check of a change: the whole GPU suite, the step trace, the bench line without the CPU / MLP legs
# usage (on the GPU box): bash tools/gpu_check.sh TAG [bench args]
tag=${1:-x}; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_gputest.log)"
timeout 300 python tools/step_trace.py > gpurun_out/${tag}_step_trace.md 2>&1; echo "trace rc=$?"
timeout 400 python bench.py --no-cpu-baseline --no-mlp "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("TPS %.3fM ms %.4f | PPS %.3fM ms %.4f | e2e %.0f / %.0f" % (d["value"]/1e6,d["ms_per_step"],d["pps"]["value"]/1e6,d["pps"]["ms_per_step"],d["e2e"]["value"],d["pps"]["e2e"]["value"]))
print("in-step", d["roofline"]["in_step_us"]["train"])
print(d["clocks"])
PY
