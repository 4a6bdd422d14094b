"""The reference's OWN callers, unmodified, driving the B200 drop-in on a GPU (the north star
names exactly these: modules/lseg_module.py, modules/lsegmentation_module.py, additional_utils/models.py).

The reference files come from /root/reference in the build container or from the byte-for-byte copy that
oracle/make_ref.sh installs under baseline/_ref/ (git-ignored, travels with the gpurun snapshot); the test skips when
neither exists. INTEGRATION.md's one-line import swap is applied without editing the file (the module name
`modules.models.lseg_net` is pre-bound to lseg_b200's LSegNet, oracle/ref_standins.py::build_reference_module).
"""
import os

import pytest
import torch

from parity_util import oracle_forward, oracle_threads, rel_err, state_dict, synth
from test_model_gpu import FULL_LOGIT_REL, _report, logit_tolerance, margin_eps
from parity_util import argmax_report

pytestmark = pytest.mark.gpu
oracle_threads()

from oracle import ref_standins as R  # noqa: E402

needs_ref = pytest.mark.skipif(not R.reference_available(), reason="reference tree not installed (oracle/make_ref.sh)")


@pytest.fixture(scope="module")
def module():
    m = R.build_reference_module(state_dict(0), drop_in=True)
    assert type(m).__module__ == "modules.lseg_module"            # the reference's own Lightning module class ...
    assert type(m.net).__module__.endswith("lseg_net") and "lang-seg_b200" in type(m.net).__module__  # ... around ours
    return m.cuda()


@needs_ref
def test_lsegmodule_constructs_and_pokes_img_size(module):
    # modules/lseg_module.py:86-89 pokes this attribute after construction
    assert module.net.pretrained.model.patch_embed.img_size == (480, 480)
    assert module.base_size == 520 and module.crop_size == 480
    assert len(module.net.labels) == 150


@needs_ref
def test_evaluate_random_and_forward_match_oracle(module):
    """LSegmentationModule.evaluate_random / forward / evaluate (modules/lsegmentation_module.py:40-59)."""
    labels = ["cat", "other", "tree", "sky"]
    x = synth.make_image(2, 96, 128, seed=41)
    ref, _ = oracle_forward(x, synth.tokenize(labels))
    with torch.no_grad():
        got = module.evaluate_random(x.cuda(), labels)
    d = {"logits": rel_err(got, ref)}
    d.update(argmax_report(got, ref, margin_eps(ref)))
    _report("reference_caller_evaluate_random", d)
    assert d["logits"] <= logit_tolerance(ref, FULL_LOGIT_REL) and d["ok"], d
    ref150, _ = oracle_forward(x[:1], synth.tokenize(synth.ade20k_labels()))
    with torch.no_grad():
        a = module(x[:1].cuda())            # LightningModule.forward -> self.net(x) with the constructor's 150 labels
        b = module.evaluate(x[:1].cuda())   # -> self.net.forward(x)
    assert torch.equal(a, b)
    assert rel_err(a, ref150) <= logit_tolerance(ref150, FULL_LOGIT_REL)


@needs_ref
def test_parent_wrapper_checkpoint_load(module):
    """Lightning checkpoints prefix every key with `net.` and carry clip_pretrained.visual.* / timm head.*; loading
    through the PARENT module with strict=True must work and must re-pack the engine weights."""
    sd = synth.make_state_dict(1, with_clip_visual_stub=True)
    x = synth.make_image(1, 64, 64, seed=2).cuda()
    before = module.evaluate_random(x, ["a", "b"]).clone()
    module.load_state_dict({"net." + k: v for k, v in sd.items()}, strict=True)
    after = module.evaluate_random(x, ["a", "b"])
    assert not torch.equal(before, after), "engine kept serving the previously packed weights"
    module.load_state_dict({"net." + k: v for k, v in state_dict(0).items()}, strict=False)
    again = module.evaluate_random(x, ["a", "b"])
    assert torch.equal(before, again)


@needs_ref
def test_multi_eval_module_threads_and_replicas(module):
    """additional_utils/models.py: LSeg_MultiEvalModule.forward (sequential multi-scale / flip / sliding window) and
    parallel_forward (DataParallel.replicate + one thread per GPU, :35-53,183-248) around the drop-in; the batched
    evaluator of lseg_b200 must reproduce the reference class bit for bit on the same network."""
    from lseg_b200.evaluator import MultiScaleEvaluator
    Ev = R.load_multi_eval_module()
    scales = [0.75, 1.25]
    ev = Ev(module, scales=scales, flip=True).cuda().eval()
    labels = ["cat", "other", "tree"]
    img = synth.make_image(1, 300, 400, seed=8).cuda()
    with torch.no_grad():
        seq = ev.forward(img, labels)
        par = ev.parallel_forward([img[0]], labels)[0]
    assert seq.shape == (1, 3, 300, 400)
    assert torch.equal(seq, par)
    glue = MultiScaleEvaluator(module.net, base_size=module.base_size, crop_size=module.crop_size, scales=scales,
                               flip=True, fused=False)(img, labels)
    assert torch.equal(glue, seq), f"batched evaluator differs from the reference class: {rel_err(glue, seq):.3e}"
    fused = MultiScaleEvaluator(module.net, base_size=module.base_size, crop_size=module.crop_size, scales=scales,
                                flip=True)(img, labels)
    # gather kernels instead of the torch glue: the network inputs agree to fp32 rounding (tests/test_evaluator_gpu.py), but
    # the network quantises its input to fp16, so a 1e-7 difference flips the rounding of a few hundred input pixels per crop
    assert rel_err(fused, seq) < logit_tolerance(seq, FULL_LOGIT_REL), rel_err(fused, seq)
