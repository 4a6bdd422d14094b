#!/usr/bin/env bash
# round-2 GPU session 21 (1 GPU): ResNet-101 zero-shot trunk (LSegRNNetZS): op pieces, model parity, golden
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -k "resnet or rn101" > gpurun_out/pytest21.log 2>&1
echo "pytest exit $?"; grep -n "^E  \|passed\|failed\|Error" gpurun_out/pytest21.log | head -30 | cut -c1-600
grep rn101 gpurun_out/parity.jsonl | cut -c1-900
