#!/usr/bin/env bash
# round-2 GPU session 8 (2 GPUs): full gpu suite on rank 0's GPU, upsample micro-bench, then the gather variants
# (background expansion kernel, trunk stream priority, low-res only) at N=2, also with the N=8 expansion load on rank 0
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
CUDA_VISIBLE_DEVICES=0 timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest8.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest8.log | cut -c1-400
CUDA_VISIBLE_DEVICES=0 python tools/op_bench.py up 2>&1 | tail -2 | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 tools/gather_check.py --steps 12 > gpurun_out/gather8_r1.json 2> gpurun_out/gather8_r1.err
tail -c 2500 gpurun_out/gather8_r1.json; tail -3 gpurun_out/gather8_r1.err | cut -c1-300
timeout 900 $TR --master-port 29532 tools/gather_check.py --steps 12 --root-repeat 4 > gpurun_out/gather8_r4.json 2> gpurun_out/gather8_r4.err
tail -c 2500 gpurun_out/gather8_r4.json; tail -3 gpurun_out/gather8_r4.err | cut -c1-300
