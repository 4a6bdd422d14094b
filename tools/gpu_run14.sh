#!/usr/bin/env bash
# round-2 GPU session 14 (1 GPU): fp32 vs fp16 low-res out_conv per level: path_1 parity and step time
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for mask in 0 14 15; do
  rm -f gpurun_out/parity.jsonl
  LSEG_OUTCONV_F16=$mask timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -k "forward_vs_oracle" > gpurun_out/pytest14_$mask.log 2>&1
  echo "mask $mask pytest exit $?"; tail -2 gpurun_out/pytest14_$mask.log | cut -c1-200
  python - <<PY
import json
for l in open('gpurun_out/parity.jsonl'):
    d=json.loads(l)
    if 'path1' in d: print('$mask', d['case'], round(d['path1']*1e3,3), round(d.get('logits_teacher_forced',0)*1e3,3), round(d.get('logit_tol_tf',0)*1e3,3))
PY
  LSEG_OUTCONV_F16=$mask timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-eval > gpurun_out/bench14_$mask.json 2> gpurun_out/bench14_$mask.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench14_$mask.json').read().strip().split('\n')[-1])
print('mask $mask', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest14.log 2>&1
echo "full pytest exit $?"; tail -4 gpurun_out/pytest14.log | cut -c1-300
