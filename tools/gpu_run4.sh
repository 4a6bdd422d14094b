#!/usr/bin/env bash
# round-2 GPU session 4: full tests, op timings (224- vs 256-wide residual tiles, LayerNorm), bench lines (default, long,
# reference arm, config 5), ncu captures for profiles/r02_*
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/op_bench.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test4.log
grep -E "passed|failed" gpurun_out/test4.log | tail -2
timeout 300 python tools/op_bench.py ln gemm > gpurun_out/op_bench4_auto.log 2>&1
LSEG_GEMM_ADD_BN=256 timeout 300 python tools/op_bench.py gemm > gpurun_out/op_bench4_bn256.log 2>&1
cat gpurun_out/op_bench4_auto.log gpurun_out/op_bench4_bn256.log | cut -c1-260
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_r02.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench4.json 2> gpurun_out/bench4.err
timeout 600 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-e2e --no-eval > gpurun_out/bench4_long.json 2> gpurun_out/bench4_long.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench4_ref.json 2> gpurun_out/bench4_ref.err
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench4_cfg5.json 2> gpurun_out/bench4_cfg5.err
for f in bench4 bench4_long bench4_ref bench4_cfg5; do echo "== $f"; tail -c 1500 gpurun_out/$f.json | cut -c1-700; tail -2 gpurun_out/$f.err | cut -c1-300; done
bash tools/profile_run.sh r02 > gpurun_out/profile_run_r02.log 2>&1
tail -5 gpurun_out/profile_run_r02.log
