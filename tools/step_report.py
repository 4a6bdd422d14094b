#!/usr/bin/env python
"""Per-launch table of one forward step from `bench.py --dump-profile` (CUDA events around every launch of one step;
no PDL overlap between launches, ~10-15 us of event overhead on the short ones): labels every launch by its position in
the engine's plan (lang-seg_b200/csrc/engine.cuh build_image_plan, clip_vitl16_384) and writes a markdown table grouped by
shape — the per-shape GEMM table of profiles/.

  python tools/step_report.py gpurun_out/profile.json [gpurun_out/traffic.csv] > profiles/r02_gemm_shapes.md

With the ncu capture of `--metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:'gemm_tc2|mhsa'` of the same command
(tools/profile_run.sh step 2) the table gains the measured DRAM MB per launch (cold L2 under ncu) next to the algorithmic
MB (operands + weights + outputs, B = 8 at 480x480).
"""
import json
import sys
from collections import OrderedDict

KINDS = {0: "elementwise", 1: "gemm", 2: "mhsa", 3: "layernorm", 4: "copy"}
HOOKS = (5, 11, 17, 23)


def labels():
    out = ["patchify", "patch embed GEMM [BT,768]x[768,1024]", "assemble tokens (+pos)"]
    for i in range(24):
        out += ["LN1", "QKV GEMM K=1024 N=3072", "MHSA", "proj GEMM K=1024 N=1024 (in-place reduce-add)", "LN2",
                "fc1 GEMM K=1024 N=4096 (+GELU)", "fc2 GEMM K=4096 N=1024 (in-place reduce-add)"]
        if i in HOOKS:
            out.append("tap copy (D2D)")
    post = ["post 1x1 1024->256", "post 1x1 1024->512", "post 1x1 1024->1024", "post 1x1 1024->1024"]
    for k in range(4):
        out += ["readout split", "readout cls GEMM (per image)", "readout tok GEMM K=1024 N=1024 (+GELU)", post[k]]
        if k == 0:
            out.append("ConvT x4 as GEMM N=4096 (depth-to-space store)")
        elif k == 1:
            out.append("ConvT x2 as GEMM N=2048 (depth-to-space store)")
        elif k == 3:
            out += ["im2col 3x3 s2", "3x3 s2 conv as GEMM K=9216 N=1024"]
    out += ["layer1_rn 3x3 256->256 @120", "layer2_rn 3x3 512->256 @60", "layer3_rn 3x3 1024->256 @30",
            "layer4_rn 3x3 1024->256 @15"]
    for k, res in ((3, 15), (2, 30), (1, 60), (0, 120)):
        if k != 3:
            out += [f"refinenet{k + 1} rcu1 conv1 @{res}", f"refinenet{k + 1} rcu1 conv2 (+x, fp32+relu out) @{res}"]
        out += [f"refinenet{k + 1} rcu2 conv1 @{res}", f"refinenet{k + 1} rcu2 conv2 (+x) @{res}",
                f"refinenet{k + 1} out_conv 1x1 @{res} (before the x2 interpolation)",
                f"refinenet{k + 1} x2 interpolation -> @{2 * res}" + (" (+ next skip)" if k else " (fp16 path_1)")]
    out += ["head1 1x1 256->512 (+row sumsq) @240", "pixel x text GEMM K=512 N=150 (NCHW store)", "logits x2 upsample -> fp32 NCHW"]
    return out


def algorithmic_mb(name, B=8):
    """operand + weight + output (+ residual) bytes of one launch, MB; None where not tabulated"""
    M = B * 901
    f16, f32 = 2, 4
    t = None
    if name.startswith("QKV"):
        t = M * 1024 * f16 + 3072 * 1024 * f16 + M * 3072 * f16
    elif name.startswith("proj"):
        t = M * 1024 * f16 + 1024 * 1024 * f16 + 2 * M * 1024 * f32
    elif name.startswith("fc1"):
        t = M * 1024 * f16 + 4096 * 1024 * f16 + M * 4096 * f16
    elif name.startswith("fc2"):
        t = M * 4096 * f16 + 4096 * 1024 * f16 + 2 * M * 1024 * f32
    elif name == "MHSA":
        t = M * 3072 * f16 + M * 1024 * f16
    elif name.startswith("head1"):
        px = B * 240 * 240
        t = px * 256 * f16 + 512 * 256 * f16 + px * 512 * f16 + px * 16 * f32
    elif name.startswith("pixel x text"):
        px = B * 240 * 240
        t = px * 512 * f16 + px * 16 * f32 + 256 * 512 * f16 + px * 150 * f16
    elif "@" in name and ("3x3" in name or "rcu" in name):
        res = int(name.rsplit("@", 1)[1])
        px = B * res * res
        cin = 256
        for c in (512, 1024):
            if f"{c}->" in name:
                cin = c
        t = px * cin * f16 + 256 * 9 * cin * f16
        if "layer" in name:
            t += px * 256 * (f32 + f16)
        elif "conv1" in name:
            t += px * 256 * f16
        elif "fp32+relu" in name:
            t += px * 256 * (f32 + f32 + f16)
        else:
            t += px * 256 * (f32 + f16)
    return None if t is None else t / 1e6


def read_traffic(path, n_expected):
    import csv
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = OrderedDict()
    for r in csv.DictReader(lines):
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    rows = [d for d in per.values() if "gemm_tc2" in d["name"] or "mhsa" in d["name"]]
    rows = rows[-n_expected:]
    return [(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / 1e6 for d in rows]


def main():
    prof = json.load(open(sys.argv[1]))
    lab = labels()
    if len(lab) != len(prof):
        print(f"<!-- label count {len(lab)} != launches {len(prof)}: plan changed, labels dropped -->")
        lab = [KINDS.get(e["kind"], "?") for e in prof]
    n_tc = sum(1 for e in prof if e["kind"] in (1, 2))
    dram = read_traffic(sys.argv[2], n_tc) if len(sys.argv) > 2 else None
    if dram is not None and len(dram) != n_tc:
        print(f"<!-- traffic capture has {len(dram)} GEMM/MHSA launches, the step {n_tc}: DRAM column dropped -->")
        dram = None
    groups = OrderedDict()
    it = iter(dram) if dram else None
    for name, e in zip(lab, prof):
        g = groups.setdefault(name, {"n": 0, "ms": 0.0, "gf": 0.0, "kind": e["kind"], "dram": 0.0, "nd": 0})
        g["n"] += 1
        g["ms"] += e["ms"]
        g["gf"] += e["gflop"]
        if it is not None and e["kind"] in (1, 2):
            g["dram"] += next(it)
            g["nd"] += 1
    total = sum(e["ms"] for e in prof)
    print("| launch | count | avg µs | GFLOP each | TFLOP/s | algorithmic MB | DRAM MB (ncu, cold L2) | share of the event-timed sum |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    for name, g in groups.items():
        us = g["ms"] / g["n"] * 1e3
        gf = g["gf"] / g["n"]
        tf = f"{g['gf'] / g['ms']:.0f}" if g["gf"] > 0.5 else "–"
        alg = algorithmic_mb(name)
        alg = f"{alg:.1f}" if alg is not None else "–"
        dr = f"{g['dram'] / g['nd']:.1f}" if g["nd"] else "–"
        print(f"| {name} | {g['n']} | {us:.1f} | {gf:.2f} | {tf} | {alg} | {dr} | {100 * g['ms'] / total:.1f} % |")
    print(f"\nsum of the event-timed launches: {total:.3f} ms ({len(prof)} launches); the step itself (launches overlapped by "
          "programmatic dependent launch, no events in between) is what `bench.py` reports as ms_per_step.")
    fam = {}
    for e in prof:
        f = fam.setdefault(KINDS.get(e["kind"], "?"), [0, 0.0, 0.0])
        f[0] += 1
        f[1] += e["ms"]
        f[2] += e["gflop"]
    print("\n| family | launches | ms | TFLOP/s |\n|---|---:|---:|---:|")
    for k, (n, ms, gf) in fam.items():
        print(f"| {k} | {n} | {ms:.3f} | {gf / ms:.0f} |" if gf > 1 else f"| {k} | {n} | {ms:.3f} | – |")


if __name__ == "__main__":
    main()
