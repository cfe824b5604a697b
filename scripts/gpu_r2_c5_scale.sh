#!/bin/bash
# configs[4] in the continuous-rk4 mode, the 8M agents sharded over N GPUs (strong scaling); usage: gpu_r2_c5_scale.sh N
set -u
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
  bench.py --gpus $N --workload c5 --adjoint-mode continuous-rk4 --steps 2 --warmup 1 > gpurun_out/c5_contrk4_n$N.json 2> gpurun_out/c5_contrk4_n$N.err
echo "N=$N exit $?" >> gpurun_out/c5_contrk4_n$N.err
python -c "
import json
l=[x for x in open('gpurun_out/c5_contrk4_n$N.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print('N=$N', d['ms_per_step'], 'ms', 'days/s', d['agent_days_per_s'], 'value', d['value'], d['scaling'], d['config']['agents_per_gpu'], d['peak_mem_gb'])"
tail -n 3 gpurun_out/c5_contrk4_n$N.err
