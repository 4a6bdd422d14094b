#!/usr/bin/env python
"""Launch one stage op a few times (profiler target: `ncu --set full -k regex:<kernel> ... python tools/op_one.py mhsa 3`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lseg_b200  # noqa: E402,F401
from lseg_b200 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "mhsa"
if what == "mhsa":
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    B, N, H = 8, 901, 16
    qkv = torch.randn(B, N, 3 * H * 64, device="cuda").half()
    for _ in range(6):
        ops.mhsa(qkv, B, N, H, False, variant=variant)
elif what == "ln":
    x = torch.randn(8 * 901, 1024, device="cuda")
    g = torch.ones(1024, device="cuda")
    b = torch.zeros(1024, device="cuda")
    for _ in range(6):
        ops.layernorm(x, g, b, 1e-6)
torch.cuda.synchronize()
print("done", what)
