#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, --set full of one predict + one train step, --set full of one
# MLP training step in tensor-core mode.  usage (GPU box): bash tools/gpu_profile.sh TAG
tag=${1:-x}
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_steps2.json 2> gpurun_out/${tag}_bench_steps2.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mlp > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_fwd|conv_bwd|gemm_tc|heads_kernel|dense_bwd|rmsprop' -o gpurun_out/${tag}_step -f python tools/prof_run.py > gpurun_out/${tag}_ncu_step.log 2>&1; echo "set full step rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm3|mlp_heads_kernel|mlp_front' --launch-skip 12 -c 12 -o gpurun_out/${tag}_mlp -f python tools/mlp_step.py 1 > gpurun_out/${tag}_ncu_mlp.log 2>&1; echo "set full mlp rc=$?"
ls -la gpurun_out/${tag}_*.ncu-rep
