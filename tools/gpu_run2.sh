#!/usr/bin/env bash
# round-2 GPU session 2: tests with the reworked tolerances, MHSA timeline + ncu stall profile, full bench lines
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not test_mhsa" > gpurun_out/test2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test2.log
timeout 120 python tools/mhsa_trace_dump.py > gpurun_out/trace2.log 2>&1
for v in 0 3; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:mhsa -s 3 -c 1 -f \
    -o gpurun_out/prof_mhsa_v${v}_r02 python tools/op_one.py mhsa $v > gpurun_out/ncu_mhsa_v$v.log 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_cfg5.json 2> gpurun_out/bench2_cfg5.err
tail -12 gpurun_out/test2.log
cat gpurun_out/bench2.json | cut -c1-600
cat gpurun_out/bench2_cfg5.json | cut -c1-400
tail -3 gpurun_out/bench2.err gpurun_out/bench2_cfg5.err
