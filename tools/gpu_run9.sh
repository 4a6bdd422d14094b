#!/usr/bin/env bash
# round-2 GPU session 9 (1 GPU): conv-then-interpolate decoder order, side-stream test, backbone goldens with recorded
# floors; upsample line-layout A/B; bench
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2700 python -m pytest tests -m gpu -q > gpurun_out/pytest9.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest9.log | cut -c1-400
python tools/op_bench.py up 2>&1 | tail -3 | cut -c1-330
timeout 900 python bench.py --steps 20 --warmup 5 --dump-profile gpurun_out/profile_r02b.json > gpurun_out/bench9b.json 2> gpurun_out/bench9b.err
python -c "
import json
d=json.loads(open('gpurun_out/bench9b.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_mhsa']['frac'], d['step_breakdown_ms'], d['clocks'], d['evaluator']['ms_per_image'])" || tail -3 gpurun_out/bench9b.err | cut -c1-300
LSEG_UPSAMPLE_SPLIT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-eval > gpurun_out/bench9b_split.json 2> gpurun_out/bench9b_split.err
python -c "
import json
d=json.loads(open('gpurun_out/bench9b_split.json').read().strip().split('\n')[-1])
print('split', d['value'], d['ms_per_step'])"
