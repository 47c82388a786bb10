#!/bin/bash
# A/B of an environment switch in the data-parallel bench.  usage: bash tools/dp_ab.sh N TAG VAR v1 v2 ...
n=$1; tag=$2; var=$3; shift 3
for v in "$@"; do
  env $var=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --no-cpu-baseline --no-mlp > gpurun_out/${tag}_n${n}_${var}${v}.json 2> gpurun_out/${tag}_n${n}_${var}${v}.err
  echo "$var=$v rc=$? $(python -c "
import json;d=json.load(open('gpurun_out/${tag}_n${n}_${var}${v}.json'));print('ms/step %.4f value %.2fM dp_check ok=%s e2e %.0f' % (d['ms_per_step'], d['value']/1e6, d.get('dp_check',{}).get('ok'), d['e2e']['value']))")"
done
