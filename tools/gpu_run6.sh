#!/usr/bin/env bash
# round-2 GPU session 6 (2 GPUs): peer-memory logits gather — bit-identity + timing of the three modes, bench lines at N=2
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/gather_check.py --steps 12 > gpurun_out/gather2.json 2> gpurun_out/gather2.err
tail -c 1500 gpurun_out/gather2.json; tail -5 gpurun_out/gather2.err | cut -c1-300
i=0
for mode in p2p_copy p2p_store nccl; do
  i=$((i+1))
  LSEG_GATHER_MODE=$mode timeout 600 $TR --master-port $((29520+i)) bench.py --gpus 2 --steps 20 --warmup 5 --no-eval \
    > gpurun_out/bench6_n2_$mode.json 2> gpurun_out/bench6_n2_$mode.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench6_n2_$mode.json').read().strip().split('\n')[-1])
print('$mode', d['value'], d['ms_per_step'], d['gather']['mode'], d['gather']['compute_only']['value'], d['gather']['root_shard_bit_identical_to_plain_forward'], d['gather']['watchdog'], d['e2e']['value'] if d.get('e2e') else None)" || tail -3 gpurun_out/bench6_n2_$mode.err | cut -c1-300
done
