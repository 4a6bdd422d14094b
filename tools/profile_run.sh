#!/usr/bin/env bash
# ncu captures behind profiles/ (run on the GPU box: gpurun -- 'bash tools/profile_run.sh r01').
# Nothing printed by a run under ncu is a bench value; bench numbers come from plain bench.py runs.
set -u
R=${1:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-eval"
# (1) launch list of the bench command: per-launch durations, cold-cache and serialised -> compare SHARES
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/launches_${R}.csv $BENCH > gpurun_out/ncu_launches_${R}.log 2>&1
# (2) DRAM traffic of every GEMM / attention launch (algorithmic-vs-actual bytes for roofline.traffic)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'gemm_tc2|mhsa' -c 1500 --csv --log-file gpurun_out/traffic_${R}.csv $BENCH > gpurun_out/ncu_traffic_${R}.log 2>&1
# (3) full-set captures: 8 consecutive GEMM launches inside the ViT trunk of a warm forward, one attention launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 --launch-skip 200 -c 8 \
  -f -o gpurun_out/prof_gemm_${R} $BENCH > gpurun_out/ncu_gemm_${R}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mhsa --launch-skip 40 -c 1 \
  -f -o gpurun_out/prof_mhsa_${R} $BENCH > gpurun_out/ncu_mhsa_${R}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:upsample2x_nchw --launch-skip 2 -c 1 \
  -f -o gpurun_out/prof_upsample_${R} $BENCH > gpurun_out/ncu_upsample_${R}.log 2>&1
# the ViT instance (fp32 residual stream, C = 1024): match the demangled name, the text tower's <__half, 512> launches come first
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'layernorm_kernel<float' --launch-skip 20 -c 1 \
  -f -o gpurun_out/prof_ln_${R} $BENCH > gpurun_out/ncu_ln_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'splitk_reduce|upsample2x_nhwc256' --launch-skip 14 -c 4 \
  -f -o gpurun_out/prof_decoder_${R} $BENCH > gpurun_out/ncu_decoder_${R}.log 2>&1
ls -la gpurun_out | tail -12
