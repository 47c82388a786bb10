#!/bin/bash
# data-parallel bench at N ranks for a list of GA3C_DP_SIDE_CTAS values.  usage: bash tools/dp_sweep.sh N TAG v1 v2 ...
n=$1; tag=$2; shift 2
for v in "$@"; do
  GA3C_DP_SIDE_CTAS=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --no-cpu-baseline --no-mlp > gpurun_out/${tag}_n${n}_side${v}.json 2> gpurun_out/${tag}_n${n}_side${v}.err
  echo "side_ctas=$v rc=$? $(python -c "
import json;d=json.load(open('gpurun_out/${tag}_n${n}_side${v}.json'));print('ms/step %.4f value %.2fM dp_check %s' % (d['ms_per_step'], d['value']/1e6, d.get('dp_check')))")"
done
