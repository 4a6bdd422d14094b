#!/usr/bin/env bash
# round-2 GPU session 23 (1 GPU, the last 4 GPU-minutes): step time of the ResNet-101 zero-shot model and one
# bench.py line with the final build (ABI v4).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T0=$SECONDS
timeout 80 python tools/rn_bench.py > gpurun_out/rn_bench.json 2> gpurun_out/rn_bench.err
echo "rn_bench exit $? at $((SECONDS - T0)) s"; cut -c1-1500 gpurun_out/rn_bench.json; tail -3 gpurun_out/rn_bench.err
LEFT=$((225 - (SECONDS - T0)))
timeout "$LEFT" python bench.py --steps 20 --warmup 5 --no-eval > gpurun_out/bench23.json 2> gpurun_out/bench23.err
echo "bench exit $? at $((SECONDS - T0)) s"; cut -c1-1200 gpurun_out/bench23.json; tail -3 gpurun_out/bench23.err
