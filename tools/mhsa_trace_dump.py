#!/usr/bin/env python
"""Raw timeline dump of the two-stream attention kernel (lseg_mhsa_trace): every clock64 stamp of the sampled CTAs as
JSON (gpurun_out/mhsa_trace.json) for offline analysis. Tags: see tools/mhsa_trace.py; 40 = first S chunk in registers,
41 = row max done, 42 = first 32 exponentials done, 43 = second S chunk in registers."""
import json
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lseg_b200  # noqa
from lseg_b200 import ops

B, N, H = 8, 901, 16
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").half()
for _ in range(3):
    ops.mhsa(qkv, B, N, H, False, variant=0)
torch.cuda.synchronize()
out, tr = ops.mhsa_trace(qkv, B, N, H, False)
torch.cuda.synchronize()
tr = tr.cpu().numpy()
dump = {}
for c in range(16):
    for w in range(10):
        e = [(int(x) >> 8, int(x) & 255) for x in tr[c, w] if x]
        if e:
            dump[f"{c}:{w}"] = e
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "mhsa_trace.json"), "w") as f:
    json.dump(dump, f)
print("stamps:", sum(len(v) for v in dump.values()))
