#!/usr/bin/env python
"""bench.py — images/sec of the LSeg forward path (480x480, ADE20K-150) on 1..8 B200.

One "step" = one LSegNet.forward over one batch of synthetic images (BASELINE.json configs[1]:
batch 8 per GPU, K=150, 480x480, ViT-L/16 DPT). Contract: see the task prompt / DESIGN.md section
"Measurement". Prints ONE JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NET_KW = dict(backbone="clip_vitl16_384", features=256, crop_size=480, arch_option=0, block_depth=0,
              activation="lrelu")
METRIC = "images/sec at 480x480 ADE20K-150 forward"


def ade_labels():
    labels = []
    with open(os.path.join(ROOT, "tests", "golden", "ade20k_objectInfo150.txt")) as f:
        for line in f.readlines():
            labels.append(line.strip().split(",")[-1].split(";")[0])
    return labels[1:]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", 1428.7), p.get("hbm_gbs", 6564.2), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.samples)}


def cpu_oracle_rate(steps, warmup, labels, size):
    """images/sec of the reference's CPU path (oracle port of lseg_net.py:160-205), B=1 per step."""
    from oracle import lseg_oracle as O
    from oracle import synth
    cores = len(os.sched_getaffinity(0))
    sd = synth.make_state_dict(0)
    tw = O.clip_text_weights_fp16(sd)
    tokens = synth.tokenize(labels)
    # torch's CPU kernels do not scale to very wide hosts on these small matrices: calibrate the thread
    # count on a reduced image and use the fastest ("all the host threads it can use")
    best, best_t = cores, None
    xs = synth.make_image(1, 160, 160, seed=0)
    for n in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(n)
        O.lseg_forward(xs, tokens[:8], sd, tw)
        t0 = time.perf_counter()
        O.lseg_forward(xs, tokens[:8], sd, tw)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    cores = best
    torch.set_num_threads(cores)
    x = synth.make_image(1, size, size, seed=0)
    for _ in range(warmup):
        O.lseg_forward(x, tokens, sd, tw)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.lseg_forward(x, tokens, sd, tw)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    labels = ade_labels()[: args.labels]
    steps = max(1, args.steps)
    rate, sec, cores = cpu_oracle_rate(steps, max(1, min(args.warmup, 2)), labels, args.size)
    sample = f"B=1 {args.size}x{args.size} K={len(labels)} forward per step, {steps} steps, text tower re-run each call"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32 trunk / fp16 text+corr", "data": "synthetic",
        "config": {"workload": f"ADE20K-150, ViT-L/16 DPT, {args.size}x{args.size}, reference CPU path (oracle port), "
                               f"batch 1 per step"},
        "cpu_baseline": {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--labels", type=int, default=150)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dump-profile", default=None, help="write the per-launch profile of one step to this JSON file")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    import lseg_b200  # noqa: F401
    from lseg_b200.lseg_net import LSegNet
    from lseg_b200.tokenizer import tokenize

    labels = ade_labels()[: args.labels]
    K = len(labels)
    B, S = args.batch, args.size
    torch.manual_seed(1234 + rank)
    net = LSegNet(labels=labels, **NET_KW).eval().to(dev)  # random-init weights of the architecture
    tokens = tokenize(labels)
    x_host = torch.randn(B, 3, S, S).clamp_(-1, 1).pin_memory()
    x = x_host.to(dev)
    eng = net._engine_for(dev)
    text = net._text_features(eng, tokens)  # encoded once, cached (steady state of a fixed label set)
    out = torch.empty((B, K, S, S), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    W, Ksteps = max(3, args.warmup), max(1, args.steps)
    for _ in range(W):
        eng.forward(x, text, K, out=out)
    barrier()
    launches_per_step = eng.last_launch_count()

    # ---- device-resident timing: K steps, L2 flushed (untimed) between steps, CUDA events per step ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(Ksteps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(Ksteps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(Ksteps):
        flush.zero_()
        starts[i].record()
        eng.forward(x, text, K, out=out)
        ends[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.summary()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = sum(step_ms)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * B * Ksteps / (total_ms / 1e3)

    # ---- multi-GPU: the one collective of the path, an all-gather of the logits ----
    gather = None
    if dist is not None:
        gathered = torch.empty((world * B, K, S, S), dtype=torch.float32, device=dev)
        comm = torch.cuda.Stream(device=dev)
        outs = [out, torch.empty_like(out)]
        for _ in range(2):
            dist.all_gather_into_tensor(gathered, out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done = [None, None]
        for i in range(Ksteps):
            o = outs[i & 1]
            if done[i & 1] is not None:
                torch.cuda.current_stream().wait_event(done[i & 1])
            eng.forward(x, text, K, out=o)
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                dist.all_gather_into_tensor(gathered, o)
                ev = torch.cuda.Event()
                ev.record(comm)
                done[i & 1] = ev
        torch.cuda.current_stream().wait_stream(comm)
        e1.record()
        barrier()
        g_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(g_ms, op=dist.ReduceOp.MAX)
        gather = {"value": world * B * Ksteps / (float(g_ms.item()) / 1e3), "unit": "images/sec",
                  "what": "forward + NCCL all-gather of fp32 logits to every rank, gather overlapped with the next step",
                  "bytes_per_rank_per_step": out.numel() * 4}
        del gathered

    # ---- roofline of the dominant kernel (tcgen05 GEMM) + MHSA: per-launch CUDA events, one step ----
    peak_tf, peak_gbs, peak_src = measured_peaks()
    roofline, mhsa_roof, breakdown = None, None, None
    if rank == 0:
        eng.forward_profiled(x, text, K, out=out)
        _, prof = eng.forward_profiled(x, text, K, out=out)
        if args.dump_profile:
            with open(args.dump_profile, "w") as f:
                json.dump([{"ms": m, "kind": k, "gflop": fl / 1e9} for m, k, fl in prof], f)
        by = {k: [0.0, 0.0, 0] for k in range(5)}  # 0 elementwise, 1 GEMM, 2 MHSA, 3 LayerNorm, 4 memset
        for ms, kind, fl in prof:
            by[kind][0] += ms
            by[kind][1] += fl
            by[kind][2] += 1
        tot = sum(v[0] for v in by.values())
        g_ms_, g_fl, g_n = by[1]
        m_ms, m_fl, m_n = by[2]
        ach = g_fl / (g_ms_ / 1e3) / 1e12 if g_ms_ > 0 else 0.0
        traffic, traffic_m = None, None  # DRAM bytes per launch from the committed ncu capture (profiles/), if any
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic.json")) as f:
                tj = json.load(f)
            traffic, traffic_m = tj["gemm"]["dram_bytes_per_launch"], tj["mhsa"]["dram_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            pass
        roofline = {"kernel": "gemm_tc2_kernel (tcgen05 GEMM / implicit conv, all launches of one step)",
                    "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "peak_source": f"{peak_src} bf16 sustained (kernel timed inside a long step)",
                    "launches": g_n, "avg_launch_ms": g_ms_ / max(g_n, 1), "share_of_step": g_ms_ / tot,
                    "traffic": traffic,
                    "traffic_note": "dram__bytes_read+write per launch, ncu capture profiles/r01_traffic.md"}
        ach_m = m_fl / (m_ms / 1e3) / 1e12 if m_ms > 0 else 0.0
        mhsa_roof = {"kernel": "mhsa2_kernel", "bound": "tensor", "achieved": ach_m, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": ach_m / peak_tf, "launches": m_n, "traffic": traffic_m,
                     "avg_launch_ms": m_ms / max(m_n, 1), "share_of_step": m_ms / tot}
        breakdown = {"gemm_ms": g_ms_, "mhsa_ms": m_ms, "layernorm_ms": by[3][0], "elementwise_ms": by[0][0],
                     "sum_ms": tot}

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        out_host = [torch.empty((B, K, S, S), dtype=torch.float32).pin_memory() for _ in range(2)]
        copy = torch.cuda.Stream(device=dev)

        def e2e_steps(n):
            for i in range(n):
                xd = x_host.to(dev, non_blocking=True)
                y = net(xd)  # public API: LSegNet.forward(x) with the constructor's labels
                ready = torch.cuda.Event()
                ready.record()
                with torch.cuda.stream(copy):
                    copy.wait_event(ready)
                    out_host[i & 1].copy_(y, non_blocking=True)
                    y.record_stream(copy)
            torch.cuda.current_stream().wait_stream(copy)

        e2e_steps(2)
        barrier()
        t0 = time.perf_counter()
        e2e_steps(Ksteps)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * Ksteps / dt, "unit": "images/sec", "h2d_bytes_per_step": x_host.numel() * 4,
               "d2h_bytes_per_step": out.numel() * 4,
               "what": "LSegNet.forward from pinned host images to pinned host fp32 logits; D2H of step i overlaps step i+1"}

    # ---- SURVEY 8(f) row 2, reported beside (not instead of) the contract's e2e: the fused argmax path returns the
    # int64 class mask every caller of the reference derives from the logits, so 8 B/pixel cross PCIe, not 4*K ----
    e2e_argmax = None
    if not args.no_e2e:
        mask_host = [torch.empty((B, S, S), dtype=torch.int64).pin_memory() for _ in range(2)]
        copy2 = torch.cuda.Stream(device=dev)

        def argmax_steps(n):
            for i in range(n):
                xd = x_host.to(dev, non_blocking=True)
                m = net.predict(xd)
                ready = torch.cuda.Event()
                ready.record()
                with torch.cuda.stream(copy2):
                    copy2.wait_event(ready)
                    mask_host[i & 1].copy_(m, non_blocking=True)
                    m.record_stream(copy2)
            torch.cuda.current_stream().wait_stream(copy2)

        argmax_steps(2)
        barrier()
        t0 = time.perf_counter()
        argmax_steps(Ksteps)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_argmax = {"value": world * B * Ksteps / dt, "unit": "images/sec", "h2d_bytes_per_step": x_host.numel() * 4,
                      "d2h_bytes_per_step": B * S * S * 8,
                      "what": "LSegNet.predict (forward fused with torch.max(.,1)[1]) from pinned host images to a pinned "
                              "host int64 mask"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, sec, cores = cpu_oracle_rate(3, 1, labels, S)
        cpu_baseline = {"value": rate, "unit": "images/sec", "cores": cores, "kind": "port",
                        "sample": f"3 timed B=1 {S}x{S} K={K} forwards after 1 warm-up ({sec:.2f} s each), oracle port of "
                                  f"lseg_net.py:160-205, fp32 trunk, text tower re-run per call"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": Ksteps, "warmup": W,
            "ms_per_step": total_ms / Ksteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands, fp32 accumulate (fp32 residual stream)", "data": "synthetic",
            "config": {"workload": f"ADE20K-150 (K={K}), ViT-L/16 DPT, {S}x{S}, batch {B} per GPU x {world} GPU",
                       "global_batch": B * world, "parallelism": f"dp{world} (batch shard, no collective on the image path)",
                       "weights": "random init of the architecture", "text_features": "cached per label set",
                       "l2": "256 MiB flush (untimed) between timed steps; per-step workspace ~3 GB >> L2"},
            "wall_s": t_wall, "clocks": clocks, "gpu_launches": launches_per_step * Ksteps,
            "launches_per_step": launches_per_step,
            "roofline": roofline, "roofline_mhsa": mhsa_roof, "step_breakdown_ms": breakdown,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "e2e_argmax": e2e_argmax,
        }
        if gather is not None:
            line["with_logits_gather"] = gather
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
