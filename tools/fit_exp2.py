"""Minimax-style fits of 2^f on [-0.5, 0.5] for the FMA-pipe exp2 (csrc/common.cuh exp2_poly<DEG>)."""
import numpy as np
from scipy.optimize import least_squares
f = np.cos(np.pi * (np.arange(4001) + 0.5) / 4001) * 0.5
t = np.exp2(f)
for deg in (3, 4, 5):
    c = np.polyfit(f, t, deg)[::-1]
    w = np.ones_like(f)
    best = None
    for it in range(300):
        r = least_squares(lambda c: (np.polyval(c[::-1], f) / t - 1) * w, c, xtol=1e-15, ftol=1e-15, gtol=1e-15)
        c = r.x
        e = np.abs(np.polyval(c[::-1], f) / t - 1)
        if best is None or e.max() < best[0]:
            best = (e.max(), c.copy())
        w *= (1 + e / e.max())
        w /= w.mean()
    emax, c = best
    ff = np.linspace(-0.5, 0.5, 200001).astype(np.float32)
    p = np.full_like(ff, np.float32(c[-1]))
    for ck in c[-2::-1]:
        p = p * ff + np.float32(ck)
    e32 = np.abs(p.astype(np.float64) / np.exp2(ff.astype(np.float64)) - 1).max()
    print(deg, "max rel err f64 %.3e  f32 eval %.3e" % (emax, e32), ["%.9ef" % v for v in c])
