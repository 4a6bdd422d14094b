#!/usr/bin/env python
"""Timeline of the CTA-pair GEMM epilogue / MMA warps (lseg_debug_gemm_trace) on the ViT GEMM shapes."""
import collections
import ctypes as C
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lseg_b200  # noqa
from lseg_b200 import ops
from lseg_b200._lib import load

TAGS = {0: "start", 1: "acc ready", 2: "tmem ld", 3: "math", 4: "buf free", 5: "sts", 6: "tma issued", 7: "tile done",
        10: "tmem_empty ok", 12: "tile issued"}
M = 7208
which = sys.argv[1] if len(sys.argv) > 1 else "fc1"
n, k = {"qkv": (3072, 1024), "proj": (1024, 1024), "fc1": (4096, 1024), "fc2": (1024, 4096)}[which]
a = torch.randn(M, k, device="cuda").half()
w = (torch.randn(n, k, device="cuda") * 0.02).half()
bias = torch.zeros(n, device="cuda")
if which in ("proj", "fc2"):
    x = torch.randn(M, n, device="cuda")
    fn = lambda: ops.gemm(a, w, n, bias=bias, res_f32=x, out_f32=x)  # noqa
else:
    o = torch.empty(M, n, device="cuda", dtype=torch.float16)
    fn = lambda: ops.gemm(a, w, n, bias=bias, act=ops.ACT_GELU if which == "fc1" else ops.ACT_NONE, out_f16=o)  # noqa
for _ in range(3):
    fn()
torch.cuda.synchronize()
tr = torch.zeros((2, 20, 512), dtype=torch.int64, device="cuda")
load().lseg_debug_gemm_trace(C.c_void_p(tr.data_ptr()))
fn()
torch.cuda.synchronize()
load().lseg_debug_gemm_trace(None)
tr = tr.cpu().numpy()
for slot in range(2):
    ts = [int(x) >> 8 for wv in range(20) for x in tr[slot, wv] if x]
    if not ts:
        continue
    t0 = min(ts)
    print(f"== pair slot {slot}: span {max(ts) - t0} clk")
    for wv in (1, 4, 8):
        e = [(int(x) >> 8, int(x) & 255) for x in tr[slot, wv] if x]
        d = collections.defaultdict(list)
        for (ta, ga), (tb, gb) in zip(e, e[1:]):
            d[(ga, gb)].append(tb - ta)
        print(f"  warp {wv}: " + " | ".join(f"{TAGS.get(x, x)}>{TAGS.get(y, y)} n={len(v)} mean={sum(v) / len(v):.0f}"
                                            for (x, y), v in d.items()))
        if slot == 0:
            print("    raw: " + " ".join(f"{t - t0}:{g}" for t, g in e[:90]))
