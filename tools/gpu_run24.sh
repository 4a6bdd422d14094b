#!/usr/bin/env bash
# round-2 GPU session 24 (1 GPU, the last 2 GPU-minutes): ncu launch list of the ResNet-101 zero-shot model's step
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 105 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/rn_launches.csv \
  python tools/rn_bench.py --steps 1 --warmup 3 > gpurun_out/rn_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/rn_launches.csv
