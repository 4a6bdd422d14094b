#!/usr/bin/env bash
# round-2 GPU session 22 (1 GPU, the round's last 8 GPU-minutes): ResNet-101 zero-shot trunk (LSegRNNetZS) op pieces +
# model parity + golden first, then as much of the rest of the suite as fits (verbose log, so a cut-off run still tells
# which tests passed). Everything under own timeouts, 400 s in total.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
T0=$SECONDS
timeout 230 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -v -k "resnet or rn101" --durations=8 \
  > gpurun_out/pytest22_rn.log 2>&1
echo "rn exit $? at $((SECONDS - T0)) s"
grep -n "PASSED\|FAILED\|ERROR\|^E  \|passed\|failed" gpurun_out/pytest22_rn.log | head -40 | cut -c1-400
grep rn101 gpurun_out/parity.jsonl | cut -c1-1200
LEFT=$((400 - (SECONDS - T0)))
if [ "$LEFT" -gt 20 ]; then
  timeout "$LEFT" python -m pytest tests/test_abi.py tests/test_ops_gpu.py tests/test_reference_callers_gpu.py \
    tests/test_evaluator_gpu.py tests/test_model_gpu.py -m gpu -v -k "not (resnet or rn101)" --durations=15 \
    > gpurun_out/pytest22_rest.log 2>&1
  echo "rest exit $? at $((SECONDS - T0)) s"
  grep -c PASSED gpurun_out/pytest22_rest.log
  grep -n "FAILED\|ERROR\|^E  \|passed\|failed" gpurun_out/pytest22_rest.log | head -30 | cut -c1-400
fi
