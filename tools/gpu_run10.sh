#!/usr/bin/env bash
# round-2 GPU session 10 (1 GPU): deterministic split-K of the low-resolution convs, fp16 low-res out_conv; A/B benches
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2700 python -m pytest tests -m gpu -q > gpurun_out/pytest10.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest10.log | cut -c1-400
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-eval \
    --dump-profile gpurun_out/profile10_$name.json > gpurun_out/bench10_$name.json 2> gpurun_out/bench10_$name.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench10_$name.json').read().strip().split('\n')[-1])
print('$name', d['value'], d['ms_per_step'], d['roofline']['frac'], d['step_breakdown_ms'], d['clocks']['sm_mhz'])" || tail -3 gpurun_out/bench10_$name.err | cut -c1-300
}
run split X=1
run nosplit LSEG_SPLITK_FIXED=0


