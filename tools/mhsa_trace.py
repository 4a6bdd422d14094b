#!/usr/bin/env python
"""Timeline of the two-stream attention kernel: clock64 stamps of sampled CTAs (lseg_mhsa_trace).

Prints, per sampled CTA, the lifetime, and per role the mean cycles between consecutive hand-off tags."""
import collections
import json
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lseg_b200  # noqa
from lseg_b200 import ops

TAGS = {0: "start", 1: "s_full", 2: "S->regs", 3: "max+exp", 4: "o_done", 5: "P stored",
        6: "p_full arr", 7: "last o_done", 8: "merged", 99: "exit", 10: "S ready",
        11: "S issued", 12: "PV ready", 13: "PV issued", 20: "k_empty", 21: "v_empty"}

B, N, H = 8, 901, 16
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").half()
for _ in range(3):
    ops.mhsa(qkv, B, N, H, False)
torch.cuda.synchronize()
out, tr = ops.mhsa_trace(qkv, B, N, H, False)
torch.cuda.synchronize()
tr = tr.cpu().numpy()
t_all0 = min(int(tr[c, w, 0]) >> 8 for c in range(16) for w in range(10) if tr[c, w, 0])
res = {}
for c in (1,):
    ev = {}
    for w in range(10):
        e = [(int(x) >> 8, int(x) & 255) for x in tr[c, w] if x]
        ev[w] = e
    if not ev[0]:
        continue
    t0 = min(e[0][0] for e in ev.values() if e)
    t1 = max(e[-1][0] for e in ev.values() if e)
    print(f"== CTA slot {c} ({'q_tile 0' if c < 8 else 'q_tile 7'}): start +{t0 - t_all0} clk, lifetime {t1 - t0} clk")
    for w, name in [(0, "tma"), (1, "mma"), (2, "softmax A q2"), (4, "softmax A q0"), (6, "softmax B q2")]:
        e = ev[w]
        d = collections.defaultdict(list)
        for (ta, ga), (tb, gb) in zip(e, e[1:]):
            d[(ga, gb)].append(tb - ta)
        parts = [f"{TAGS.get(a, a)}>{TAGS.get(b_, b_)} {sum(v) / len(v):.0f}" for (a, b_), v in d.items() if len(v) > 2]
        print(f"  [{name}] " + " | ".join(parts))
    if c == 1:
        for w in range(3):
            print(f"  raw warp {w}: " + " ".join(f"{t - t0}:{g}" for t, g in ev[w] if t - t0 < 20000))
lifetimes = []
for c in range(16):
    ts = [int(x) >> 8 for w in range(10) for x in tr[c, w] if x]
    if ts:
        lifetimes.append(max(ts) - min(ts))
print("CTA lifetimes (clk):", lifetimes)
