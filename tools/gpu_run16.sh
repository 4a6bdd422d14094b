#!/usr/bin/env bash
# round-2 GPU session 16 (2 GPUs): pipelined gather (expansion of step s-1 on the gathering rank's main stream)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tools/gather_check.py --steps 12 > gpurun_out/gather16_r1.json 2> gpurun_out/gather16_r1.err
python -c "
import json
d=json.loads(open('gpurun_out/gather16_r1.json').read().strip().split('\n')[-1])
print({k:(round(v['ms_per_step'],3), v.get('bit_identical')) for k,v in d.items() if isinstance(v,dict)})" || tail -5 gpurun_out/gather16_r1.err | cut -c1-300
timeout 600 $TR --master-port 29542 tools/gather_check.py --steps 12 --root-repeat 4 > gpurun_out/gather16_r4.json 2> gpurun_out/gather16_r4.err
python -c "
import json
d=json.loads(open('gpurun_out/gather16_r4.json').read().strip().split('\n')[-1])
print({k:(round(v['ms_per_step'],3), v.get('bit_identical')) for k,v in d.items() if isinstance(v,dict)})" || tail -5 gpurun_out/gather16_r4.err | cut -c1-300
for flag in "" "--gather-sidestream"; do
LSEG_BENCH_ROOT_BATCH=7 timeout 600 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --no-eval $flag > gpurun_out/bench16_n2$flag.json 2> gpurun_out/bench16_n2$flag.err
python -c "
import json
d=json.loads(open('gpurun_out/bench16_n2$flag.json').read().strip().split('\n')[-1])
g=d['gather']
print('N=2 $flag', d['value'], d['ms_per_step'], 'lowres', g['lowres_only']['value'], 'balanced', (g.get('balanced') or {}).get('value'), 'compute', g['compute_only']['value'], g['root_shard_bit_identical_to_plain_forward'], g['watchdog'])" || tail -5 "gpurun_out/bench16_n2$flag.err" | cut -c1-400
done
