#!/usr/bin/env bash
# round-2 GPU session 13 (8 GPUs): the scaling lines with the peer-memory gather inside the timed step
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
for n in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n))"
  timeout 900 $TR bench.py --gpus $n --steps 20 --warmup 5 --no-eval > gpurun_out/bench13_n$n.json 2> gpurun_out/bench13_n$n.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench13_n$n.json').read().strip().split('\n')[-1])
g=d['gather']
print('N=$n', d['value'], d['ms_per_step'], 'lowres', g['lowres_only']['value'], 'balanced', g.get('balanced'), 'compute', g['compute_only']['value'], g['root_shard_bit_identical_to_plain_forward'], g['watchdog'], 'e2e', d['e2e']['value'] if d.get('e2e') else None)" || tail -5 gpurun_out/bench13_n$n.err | cut -c1-400
done
