"""bench.py host logic that runs without a GPU: both arms (`--impl b200`, `--impl reference`) describe ONE workload — the
`config` object of their JSON lines is the same dict, so the driver's ratio compares like with like."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def _args(**kw):
    base = dict(gpus=1, steps=20, warmup=5, impl="b200", config=2, batch=0, labels=0, size=480, gather="p2p_copy",
                backbone="clip_vitl16_384")
    base.update(kw)
    return argparse.Namespace(**base)


def test_both_arms_print_the_same_config():
    for world, cfg_id in ((1, 2), (8, 2), (1, 5)):
        a, r = _args(gpus=world, config=cfg_id), _args(gpus=world, config=cfg_id, impl="reference")
        ca, cr = bench.line_config(bench.make_config(a), a, world), bench.line_config(bench.make_config(r), r, world)
        assert ca == cr
        json.dumps(ca)
        assert ca["global_batch"] == ca["batch_per_gpu"] * world
        assert "flush" in ca["l2"] and "workload" in ca
    c = bench.line_config(bench.make_config(_args()), _args(), 1)
    assert c["batch_per_gpu"] == 8 and "K=150" in c["workload"] and "480x480" in c["workload"]
