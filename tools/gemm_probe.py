#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM on the LSeg shapes, with the measurement-only probe modes of
GemmParams::probe (LSEG_GEMM_PROBE: 1 skip epilogue, 2 skip TMA, 4 skip MMA; results are garbage then).

  for p in 0 1 2 3 5 6; do LSEG_GEMM_PROBE=$p python tools/gemm_probe.py; done
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lseg_b200  # noqa: E402,F401
from lseg_b200 import ops  # noqa: E402


def bench(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us


def main():
    dev = "cuda"
    M = 7208
    out = {"probe": int(os.environ.get("LSEG_GEMM_PROBE", "0")), "one_cta": os.environ.get("LSEG_GEMM_1CTA", "0")}
    shapes = {"qkv": (M, 3072, 1024), "proj": (M, 1024, 1024), "fc1": (M, 4096, 1024), "fc2": (M, 1024, 4096)}
    for name, (m, n, k) in shapes.items():
        a = torch.randn(m, k, device=dev).half()
        w = (torch.randn(n, k, device=dev) * 0.02).half()
        bias = torch.zeros(n, device=dev)
        if name in ("proj", "fc2"):
            x = torch.randn(m, n, device=dev)
            fn = lambda: ops.gemm(a, w, n, bias=bias, res_f32=x, out_f32=x)  # noqa: E731
        elif name == "fc1":
            o = torch.empty(m, n, device=dev, dtype=torch.float16)
            fn = lambda: ops.gemm(a, w, n, bias=bias, act=ops.ACT_GELU, out_f16=o)  # noqa: E731
        else:
            o = torch.empty(m, n, device=dev, dtype=torch.float16)
            fn = lambda: ops.gemm(a, w, n, bias=bias, out_f16=o)  # noqa: E731
        us = bench(fn)
        out[name] = {"us": round(us, 1), "tflops": round(2 * m * n * k / us / 1e6, 1)}
    # 3x3 conv 256->256 at 120x120, B=8 (RCU conv1: scale+bias+relu -> fp16)
    xin = torch.randn(8, 120, 120, 256, device=dev).half()
    wc = (torch.randn(256, 9 * 256, device=dev) * 0.02).half()
    sc = torch.ones(256, device=dev)
    oc = torch.empty(8, 120, 120, 256, device=dev, dtype=torch.float16)
    us = bench(lambda: ops.gemm(xin, wc, 256, conv=(3, 1), scale=sc, bias=sc, act=ops.ACT_RELU, out_f16=oc, ldc=256))
    out["conv3x3"] = {"us": round(us, 1), "tflops": round(2 * 8 * 120 * 120 * 256 * 9 * 256 / us / 1e6, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
