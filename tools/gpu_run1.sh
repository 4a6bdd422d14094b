#!/usr/bin/env bash
# round-2 GPU session 1: full GPU test suite, op micro-benchmarks (all MHSA variants), in-step A/B of the MHSA kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test1.log
timeout 400 python tools/op_bench.py mhsa ln gemm > gpurun_out/op_bench1.log 2>&1
for v in 0 1 2 3 4; do
  LSEG_MHSA_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e \
    > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
done
tail -15 gpurun_out/test1.log
grep -h '"value"' gpurun_out/bench_v*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d.get('roofline_mhsa', {}).get('achieved'), d.get('roofline', {}).get('achieved'))"
