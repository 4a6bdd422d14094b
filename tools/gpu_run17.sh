#!/usr/bin/env bash
# round-2 GPU session 17 (1 GPU): spill-free register-direct epilogue (NCHW-T store bounds), full suite, bench + profile,
# ncu: ViT LayerNorm instance, per-launch traffic
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest17.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest17.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 --dump-profile gpurun_out/profile_r02_final.json > gpurun_out/bench17.json 2> gpurun_out/bench17.err
timeout 900 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-eval > gpurun_out/bench17_long.json 2> gpurun_out/bench17_long.err
timeout 900 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu-baseline --no-eval > gpurun_out/bench17_cfg5.json 2> gpurun_out/bench17_cfg5.err
for f in bench17 bench17_long bench17_cfg5; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().split('\n')[-1])
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_mhsa') or {}).get('frac'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('e2e') or {}).get('value'))"; done
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-eval"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'gemm_tc2|mhsa' -c 1500 --csv --log-file gpurun_out/traffic_r02.csv $BENCH > gpurun_out/ncu_traffic_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'layernorm_kernel<float' --launch-skip 20 -c 1 \
  -f -o gpurun_out/prof_ln_r02 $BENCH > gpurun_out/ncu_ln_r02.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/ncu_launches_r02.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit,clocks_throttle_reasons.active --format=csv > gpurun_out/clocks_r02.txt 2>&1
ls -la gpurun_out/prof_ln_r02.ncu-rep gpurun_out/traffic_r02.csv
