#!/usr/bin/env python
"""Multi-GPU check + timing of lseg_b200.parallel.LogitsGather (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/gather_check.py [--batch 8] [--steps 12]

For every mode (p2p_copy, p2p_store, nccl) the gathered fp32 [world*B,K,H,W] tensor on rank 0 must equal, bit for bit, the
concatenation of what each rank's plain LSegNet.forward returns (checked through an NCCL all-gather of those), several
steps in a row (double-buffer reuse), and the per-step time of the pipelined gather is reported next to compute-only."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--labels", type=int, default=150)
    ap.add_argument("--root-repeat", type=int, default=1,
                    help="experiment: rank 0 expands the gathered shards this many times per step (at world 2, 4 = the "
                         "fp32 expansion load of world 8)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import lseg_b200  # noqa: F401
    from lseg_b200 import ops, tokenizer
    from lseg_b200.lseg_net import LSegNet
    from lseg_b200.parallel import LogitsGather
    tokenizer.enable_stand_in()
    labels = [f"label{i}" for i in range(args.labels)]
    net = LSegNet(labels=labels, backbone="clip_vitl16_384", features=256, crop_size=480, arch_option=0, block_depth=0,
                  activation="lrelu").eval().to(dev)  # same seeded init on every rank
    B, K, S = args.batch, args.labels, args.size
    eng = net._engine_for(dev)
    text = net._text_features(eng, net.text)
    xs = [torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(100 * rank + i)).clamp_(-1, 1).to(dev)
          for i in range(3)]
    res = {}
    # (mode, background kernel for rank 0's fp32 expansion, trunk on a high-priority stream, materialise fp32)
    variants = [("p2p_copy", "p2p_copy", False, False, True), ("p2p_copy_pipelined", "p2p_copy", False, False, True),
                ("p2p_copy_bg", "p2p_copy", True, False, True),
                ("p2p_copy_hp", "p2p_copy", False, True, True), ("p2p_copy_lowres", "p2p_copy", False, False, False),
                ("p2p_store", "p2p_store", False, False, True), ("p2p_store_pipelined", "p2p_store", False, False, True),
                ("nccl", "nccl", False, False, True)]
    lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
    hp_stream = torch.cuda.Stream(device=dev, priority=hi)
    for name, mode, background, high_prio, materialize in variants:
        g = LogitsGather(eng, B, K, S, S, root=0, mode=mode, background=background, materialize=materialize,
                         pipelined=name.endswith("_pipelined"))
        g.repeat = args.root_repeat
        trunk = hp_stream if high_prio else torch.cuda.current_stream(dev)
        torch.cuda.synchronize()
        ok = True
        for i in range(5 if materialize else 0):  # correctness over several steps (slot reuse), different inputs per step
            x = xs[i % 3]
            with torch.cuda.stream(trunk):
                g.forward(x, text)
                full = g.flush()
            torch.cuda.synchronize()
            own = eng.forward(x, text, K)
            ref = torch.empty((world * B, K, S, S), dtype=torch.float32, device=dev) if rank == 0 else None
            # reference gather of the plain forwards (fp32, NCCL)
            lst = [torch.empty_like(own) for _ in range(world)] if rank == 0 else None
            dist.gather(own, lst, dst=0)
            if rank == 0:
                ref = torch.cat(lst, 0)
                ok = ok and bool(torch.equal(full, ref))
                del ref, lst
            del own
            torch.cuda.synchronize()  # the plain forward above shares the engine's buffers with the next gather step
        # timing: pipelined steps
        with torch.cuda.stream(trunk):
            for _ in range(3):
                g.forward(xs[0], text)
            g.flush()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(trunk):
            e0.record()
            for i in range(args.steps):
                g.forward(xs[i % 3], text)
            g.flush()
            e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wd = ops.read_watchdog()
        res[name] = {"ran_as": g.mode, "background": background, "trunk_high_priority": high_prio,
                     "materialize": materialize, "root_repeat": args.root_repeat, "fallback": g.fallback_reason, "bit_identical": ok if rank == 0 else None,
                     "ms_per_step": float(t.item()), "img_per_s": world * B / float(t.item()) * 1e3, "watchdog": wd[0]}
        g.close()
        dist.barrier()
    # compute only
    out = torch.empty((B, K, S, S), dtype=torch.float32, device=dev)
    for _ in range(3):
        eng.forward(xs[0], text, K, out=out)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        eng.forward(xs[i % 3], text, K, out=out)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["compute_only"] = {"ms_per_step": float(t.item()), "img_per_s": world * B / float(t.item()) * 1e3}
    if rank == 0:
        print(json.dumps({"world": world, "batch_per_gpu": B, "K": K, "size": S, **res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
