#!/usr/bin/env python
"""HBM write/copy bandwidth reference points for the HBM-bound kernels (upsample, LayerNorm)."""
import json
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lseg_b200  # noqa
from lseg_b200 import ops


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


out = {}
y = torch.empty(8, 150, 480, 480, device="cuda")
out["fill_1.1GB_us"] = bench(lambda: y.fill_(1.0))
out["fill_GBs"] = y.numel() * 4 / out["fill_1.1GB_us"] / 1e3
z = torch.empty_like(y)
out["copy_1.1GB_us"] = bench(lambda: z.copy_(y))
out["copy_rw_GBs"] = 2 * y.numel() * 4 / out["copy_1.1GB_us"] / 1e3
lr = torch.randn(8, 150, 240, 240, device="cuda").half()
out["upsample_us"] = bench(lambda: ops.upsample2x_nchw(lr))
out["upsample_write_GBs"] = y.numel() * 4 / out["upsample_us"] / 1e3
xh = torch.randn(8, 120, 120, 256, device="cuda").half()
out["upsample_nhwc_us"] = bench(lambda: ops.upsample2x_nhwc(xh), 20)
x = torch.randn(7208, 1024, device="cuda")
g = torch.ones(1024, device="cuda")
out["layernorm_us"] = bench(lambda: ops.layernorm(x, g, g, 1e-6), 50)
out["layernorm_rw_GBs"] = 7208 * 1024 * 6 / out["layernorm_us"] / 1e3
qkv = torch.randn(8, 901, 3072, device="cuda").half()
out["mhsa_us"] = bench(lambda: ops.mhsa(qkv, 8, 901, 16, False), 20)
out["mhsa_tflops"] = 4.0 * 8 * 16 * 901 * 901 * 64 / out["mhsa_us"] / 1e6
out["env"] = {k: v for k, v in os.environ.items() if k.startswith("LSEG_")}
print(json.dumps(out))
