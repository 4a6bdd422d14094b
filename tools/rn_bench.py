"""Step time of the zero-shot model on the ResNet-101 trunk (LSegRNNetZS, BASELINE configs[3]: B=8, 480x480, PASCAL-20
label file, K=2 per image): CUDA events per step on the launching stream, L2 flushed (untimed) between steps, random-init
weights of the architecture, inputs resident in HBM. Prints ONE JSON line: ms/step, images/s, launches per step, and the
per-kind sums of one event-profiled step (tcgen05 GEMM family: time, algorithmic TFLOP/s against MEASURED_PEAKS.json).

    python tools/rn_bench.py [--batch 8] [--size 480] [--steps 20] [--warmup 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    import lseg_b200  # noqa: F401
    from lseg_b200 import tokenizer
    from lseg_b200.lseg_net import LSegRNNetZS
    tokenizer.enable_stand_in()
    dev = torch.device("cuda", 0)
    names = [ln.strip() for ln in open(os.path.join(ROOT, "tests", "golden", "fewshot_pascal.txt")) if ln.strip()]
    torch.manual_seed(4321)
    net = LSegRNNetZS(label_list=names, features=256, arch_option=0, block_depth=0, activation="lrelu").eval().to(dev)
    B, S = a.batch, a.size
    x = torch.randn(B, 3, S, S).clamp_(-1, 1).to(dev)
    class_info = torch.randint(0, len(names), (B,))
    eng = net._engine_for(dev)
    text = net._image_text(eng, class_info, dev)  # pair features cached: the steady state of a fixed label file
    out = torch.empty((B, 2, S, S), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(3, a.warmup)):
        eng.forward(x, text, 2, text_image_stride=2, out=out)
    torch.cuda.synchronize()
    st = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    for i in range(a.steps):
        flush.zero_()
        st[i].record()
        eng.forward(x, text, 2, text_image_stride=2, out=out)
        en[i].record()
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in zip(st, en)) / a.steps
    launches = eng.last_launch_count()
    _, prof = eng.forward_profiled(x, text, 2, text_image_stride=2, out=out)
    by = {}
    for t, kind, fl in prof:
        v = by.setdefault(int(kind), [0.0, 0.0, 0])
        v[0] += t
        v[1] += fl
        v[2] += 1
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        peaks = {}
    g = by.get(1, [0.0, 0.0, 0])
    tot = sum(v[0] for v in by.values())
    line = {"what": "LSegRNNetZS.forward (ResNet-101 trunk, zero-shot head), device-resident inputs", "batch": B, "size": S,
            "steps": a.steps, "ms_per_step": ms, "images_per_sec": B / (ms / 1e3), "launches_per_step": launches,
            "profiled_step_ms": tot,
            "gemm_family": {"launches": g[2], "ms": g[0], "algorithmic_gflop": g[1] / 1e9,
                            "tflops": (g[1] / (g[0] / 1e3) / 1e12) if g[0] else None, "share_of_step": g[0] / tot if tot else None},
            "other_kinds_ms": {str(k): v[0] for k, v in by.items() if k != 1},
            "measured_peaks": peaks}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
