#!/bin/bash
# strong-scaling lines (BASELINE.json configs[3]: the same 1M agents sharded over N GPUs); usage: gpu_scale.sh "2 4 8"
set -u
mkdir -p gpurun_out
for N in $1; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  echo "N=$N exit $?" >> gpurun_out/scale_n$N.err
  python -c "
import json
l=[x for x in open('gpurun_out/scale_n$N.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print('N=$N', d['ms_per_step'], 'ms', 'days/s', d['agent_days_per_s'], 'value', d['value'], d['scaling'], d['config']['agents_per_gpu'])"
  tail -n 3 gpurun_out/scale_n$N.err
done
