#!/usr/bin/env bash
# round-2 GPU session 5: mhsa5 (two query tiles per CTA) correctness + speed
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider -k "test_mhsa" > gpurun_out/test5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test5.log
grep -E "passed|failed" gpurun_out/test5.log | tail -2
grep -E "^FAILED" gpurun_out/test5.log | head -20
timeout 400 python tools/op_bench.py mhsa > gpurun_out/op_bench5.log 2>&1
cut -c1-330 gpurun_out/op_bench5.log
for v in 0 9 10 11; do
  LSEG_MHSA_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-eval \
    > gpurun_out/bench5_v$v.json 2> gpurun_out/bench5_v$v.err
done
grep -h '"value"' gpurun_out/bench5_v*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d.get('roofline_mhsa', {}).get('achieved'), d.get('roofline', {}).get('achieved'))"
tail -2 gpurun_out/bench5_v9.err
