#!/usr/bin/env python
"""Stage-op micro-benchmarks on the GPU box (CUDA events on the launching stream, warm-up, rotating inputs so that no
launch finds its input in L2 from the previous one). One JSON line per case into gpurun_out/op_bench.jsonl and stdout.

  python tools/op_bench.py mhsa          # all lseg_mhsa_variant kernels at the bench shapes + correctness
  python tools/op_bench.py ln | gemm     # LayerNorm / the ViT GEMM shapes
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lseg_b200  # noqa: E402,F401
from lseg_b200 import ops  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "op_bench.jsonl")
PEAK_TF = 1428.7


def emit(d):
    line = json.dumps(d)
    print(line, flush=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(line + "\n")


def time_launches(fn, n_bufs, iters=40, warmup=8):
    """fn(i) launches on buffer set i % n_bufs. Returns (median_us, min_us) of single launches."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    us = [a.elapsed_time(b) * 1e3 for a, b in evs]
    return statistics.median(us), min(us)


def time_back_to_back(fn, n_bufs, iters=48, warmup=8):
    """Average over one event pair around `iters` back-to-back launches (PDL overlap included, like inside a step)."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def bench_mhsa():
    shapes = [(8, 901, 16, False, "cfg2 480^2 B=8"), (4, 2117, 16, False, "cfg5 736^2 B=4"), (150, 77, 8, True, "text K=150")]
    for B, N, heads, causal, tag in shapes:
        D = heads * 64
        nb = 5
        g = torch.Generator(device="cpu").manual_seed(5)
        qkvs = [(torch.randn((B, N, 3 * D), generator=g)).half().cuda() for _ in range(nb)]
        outs = [torch.empty((B * N, D), dtype=torch.float16, device="cuda") for _ in range(nb)]
        q, k, v = qkvs[0][:1].float().view(1, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
        s = (q @ k.transpose(-1, -2)) * 0.125
        if causal:
            s = s + torch.full((N, N), float("-inf"), device="cuda").triu_(1)
        ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(N, D)
        flops = 4.0 * B * heads * N * N * 64 * (0.5 if causal else 1.0)
        for variant in (0, 3, 6, 7):
            def fn(i, variant=variant):
                ops.mhsa(qkvs[i % nb], B, N, heads, causal, variant=variant, out=outs[i % nb])
            try:
                med, mn = time_launches(fn, nb)
                b2b = time_back_to_back(fn, nb)
                err = ((outs[0][:N].float() - ref).abs().max() / ref.abs().max()).item()
                wd = ops.read_watchdog()
                emit({"op": "mhsa", "case": tag, "variant": variant, "median_us": med, "min_us": mn, "b2b_us": b2b,
                      "tflops_median": flops / med / 1e6, "tflops_b2b": flops / b2b / 1e6,
                      "frac_sustained_b2b": flops / b2b / 1e6 / PEAK_TF, "rel_err": err, "watchdog": wd[0]})
            except Exception as e:  # keep going: the other variants are still informative
                emit({"op": "mhsa", "case": tag, "variant": variant, "error": str(e)[:300]})


def bench_ln():
    M, C = 8 * 901, 1024
    nb = 6
    xs = [torch.randn((M, C), device="cuda") for _ in range(nb)]
    g = torch.ones(C, device="cuda")
    b = torch.zeros(C, device="cuda")

    def fn(i):
        ops.layernorm(xs[i % nb], g, b, 1e-6)
    med, mn = time_launches(fn, nb)
    b2b = time_back_to_back(fn, nb)
    byts = M * C * 6
    emit({"op": "layernorm", "case": "M=7208 C=1024 fp32->fp16", "median_us": med, "min_us": mn, "b2b_us": b2b,
          "gbs_b2b": byts / b2b / 1e3, "frac_hbm_b2b": byts / b2b / 1e3 / 6564.2})


def bench_upsample():
    """fp32 logits expansion: the shared-memory kernel of the step and the background kernel of the multi-GPU gather."""
    B, K, H, W = 8, 150, 240, 240
    nb = 3
    xs = [torch.randn((B, K, H, W), device="cuda").half() for _ in range(nb)]
    byts = B * K * H * W * (2 + 16)
    lib = ops.load()
    import ctypes as C
    for name, fn, split in (("upsample2x_nchw interleaved line", lib.lseg_upsample2x_nchw, 0),
                            ("upsample2x_nchw parity-split line", lib.lseg_upsample2x_nchw, 1),
                            ("upsample2x_nchw_bg", lib.lseg_upsample2x_nchw_bg, 0)):
        lib.lseg_debug_upsample_layout(split)
        outs = [torch.empty((B, K, 2 * H, 2 * W), device="cuda") for _ in range(nb)]

        def f(i, fn=fn):
            ops.check(fn(C.c_void_p(xs[i % nb].data_ptr()), C.c_void_p(outs[i % nb].data_ptr()), B * K, H, W,
                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        med, mn = time_launches(f, nb, iters=20, warmup=4)
        b2b = time_back_to_back(f, nb, iters=20, warmup=4)
        emit({"op": name, "case": "8x150x240x240 fp16 -> 480x480 fp32", "median_us": med, "min_us": mn, "b2b_us": b2b,
              "gbs_b2b": byts / b2b / 1e3, "frac_hbm_b2b": byts / b2b / 1e3 / 6564.2})
        del outs
    lib.lseg_debug_upsample_layout(0)


def bench_gemm():
    M = 8 * 901
    cases = [("qkv", 3072, 1024, "f16"), ("proj", 1024, 1024, "add"), ("fc1", 4096, 1024, "gelu"), ("fc2", 1024, 4096, "add")]
    for name, N, K, kind in cases:
        nb = 4
        a = [ops.pad_rows((torch.randn((M, K), device="cuda") * 0.5).half()) for _ in range(nb)]
        w = ops.pad_rows((torch.randn((N, K), device="cuda") * 0.03).half())
        bias = torch.randn(N, device="cuda") * 0.1
        x32 = [torch.zeros((M, N), device="cuda") for _ in range(nb)]
        o16 = [torch.empty((M, N), dtype=torch.float16, device="cuda") for _ in range(nb)]

        def fn(i):
            j = i % nb
            if kind == "add":
                ops.gemm(a[j], w, N, M=M, bias=bias, res_f32=x32[j], out_f32=x32[j])
            elif kind == "gelu":
                ops.gemm(a[j], w, N, M=M, bias=bias, act=ops.ACT_GELU, out_f16=o16[j])
            else:
                ops.gemm(a[j], w, N, M=M, bias=bias, out_f16=o16[j])
        med, mn = time_launches(fn, nb)
        b2b = time_back_to_back(fn, nb)
        flops = 2.0 * M * N * K
        emit({"op": "gemm", "case": f"{name} M={M} N={N} K={K}", "median_us": med, "min_us": mn, "b2b_us": b2b,
              "tflops_b2b": flops / b2b / 1e6, "frac_sustained_b2b": flops / b2b / 1e6 / PEAK_TF})


if __name__ == "__main__":
    what = sys.argv[1:] or ["mhsa", "ln", "gemm", "up"]
    if "mhsa" in what:
        bench_mhsa()
    if "ln" in what:
        bench_ln()
    if "gemm" in what:
        bench_gemm()
    if "up" in what:
        bench_upsample()
