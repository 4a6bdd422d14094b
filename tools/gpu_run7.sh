#!/usr/bin/env bash
# round-2 GPU session 7 (1 GPU): new backbones (clip_vitb32_384, clipRN50x16_vitl16_384), parity-split upsample line,
# re-batched fused evaluator: the full gpu suite, then the bench lines
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest7.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest7.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
tail -c 3000 gpurun_out/bench7.json; tail -3 gpurun_out/bench7.err | cut -c1-300
for bb in clip_vitb32_384 clipRN50x16_vitl16_384; do
  timeout 600 python bench.py --steps 20 --warmup 5 --backbone $bb --no-cpu-baseline --no-eval \
    > gpurun_out/bench7_$bb.json 2> gpurun_out/bench7_$bb.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench7_$bb.json').read().strip().split('\n')[-1])
print('$bb', d['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'))" || tail -3 gpurun_out/bench7_$bb.err | cut -c1-300
done
python tools/op_bench.py ln > gpurun_out/op_ln7.log 2>&1; tail -2 gpurun_out/op_ln7.log | cut -c1-300
