#!/usr/bin/env bash
# round-2 GPU session 15 (1 GPU): the evidence behind profiles/r02_* with the final kernels — ncu captures first, then the
# plain bench lines (nothing printed under ncu is a bench value)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
bash tools/profile_run.sh r02 > gpurun_out/profile_run_r02.log 2>&1
tail -3 gpurun_out/profile_run_r02.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit,clocks_throttle_reasons.active --format=csv > gpurun_out/clocks_r02.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --dump-profile gpurun_out/profile_r02_final.json > gpurun_out/bench15.json 2> gpurun_out/bench15.err
tail -c 600 gpurun_out/bench15.json; echo
timeout 900 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-eval > gpurun_out/bench15_long.json 2> gpurun_out/bench15_long.err
timeout 900 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu-baseline --no-eval > gpurun_out/bench15_cfg5.json 2> gpurun_out/bench15_cfg5.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench15_ref.json 2> gpurun_out/bench15_ref.err
python tools/op_bench.py > gpurun_out/op_bench15.log 2>&1
for f in bench15 bench15_long bench15_cfg5 bench15_ref; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().split('\n')[-1])
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('clocks') or {}).get('sm_mhz'))"; done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
