#!/usr/bin/env bash
# round-2 GPU session 3: mhsa4 (P in TMEM) correctness + speed, arch_option blocks, reworked text tests, full bench line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test3.log
timeout 400 python tools/op_bench.py mhsa > gpurun_out/op_bench3.log 2>&1
for v in 3 6 7; do
  LSEG_MHSA_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e \
    > gpurun_out/bench3_v$v.json 2> gpurun_out/bench3_v$v.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_cfg5.json 2> gpurun_out/bench3_cfg5.err
grep -E "passed|failed" gpurun_out/test3.log | tail -3
grep -h '"value"' gpurun_out/bench3_v*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d.get('roofline_mhsa', {}).get('achieved'), d.get('roofline', {}).get('achieved'))"
cut -c1-400 gpurun_out/bench3.json; tail -2 gpurun_out/bench3.err; cut -c1-300 gpurun_out/bench3_cfg5.json; tail -2 gpurun_out/bench3_cfg5.err
