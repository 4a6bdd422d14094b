#!/usr/bin/env bash
# round-2 GPU session 20 (1 GPU): final build — full suite, smoke, bench lines, pixel x text under ncu
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest20.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest20.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --dump-profile gpurun_out/profile_r02_final.json > gpurun_out/bench20.json 2> gpurun_out/bench20.err
timeout 900 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-eval > gpurun_out/bench20_long.json 2> gpurun_out/bench20_long.err
for f in bench20 bench20_long; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().split('\n')[-1])
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_mhsa') or {}).get('frac'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('e2e') or {}).get('value'))"; done
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-eval"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'gemm_tc2|mhsa' -c 1500 --csv --log-file gpurun_out/traffic_r02.csv $BENCH > gpurun_out/ncu_traffic_r02.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/ncu_launches_r02.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit,clocks_throttle_reasons.active --format=csv > gpurun_out/clocks_r02.txt 2>&1
