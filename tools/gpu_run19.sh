#!/usr/bin/env bash
# round-2 GPU session 19 (4 GPUs): the default (pipelined) gather at N=4 with the final kernels
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
n=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29704"
timeout 600 $TR bench.py --gpus $n --steps 20 --warmup 5 --no-eval > gpurun_out/bench19_n$n.json 2> gpurun_out/bench19_n$n.err
python -c "
import json
d=json.loads(open('gpurun_out/bench19_n$n.json').read().strip().split('\n')[-1])
g=d['gather']
print('N=$n', d['value'], d['ms_per_step'], 'lowres', g['lowres_only']['value'], 'balanced', g.get('balanced'), 'compute', g['compute_only']['value'], g['root_shard_bit_identical_to_plain_forward'], g['watchdog'], 'e2e', d['e2e']['value'] if d.get('e2e') else None)" || tail -5 gpurun_out/bench19_n$n.err | cut -c1-400
