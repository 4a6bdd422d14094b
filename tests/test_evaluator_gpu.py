"""GPU tests of the multi-scale / flip / sliding-window evaluator and of the fused
preprocessing kernel (row 3): the gather kernels of csrc/evaluator.cuh against the torch-glue implementation (which
tests/test_evaluator_cpu.py pins bit for bit against the unmodified reference class) on the real LSegNet."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from parity_util import NET_KW, rel_err, state_dict, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net():
    import lseg_b200  # noqa: F401
    from lseg_b200.lseg_net import LSegNet
    n = LSegNet(labels=["cat", "other", "tree"], **NET_KW)
    n.load_state_dict(state_dict(0))
    return n.cuda().eval()


# (image h, w, base_size, crop_size, scales): whole-image scales, sliding windows in one and both directions, portrait
CASES = [(150, 200, 160, 128, (0.5, 1.0, 1.75)), (210, 140, 200, 160, (0.75, 1.25)), (96, 96, 128, 96, (0.5, 1.5, 2.25))]


@pytest.mark.parametrize("h,w,base,crop,scales", CASES)
def test_fused_matches_torch_glue_and_is_batch_invariant(net, h, w, base, crop, scales):
    from lseg_b200.evaluator import MultiScaleEvaluator
    img = synth.make_image(1, h, w, seed=h + w).cuda()
    labels = ["cat", "other", "tree"]
    kw = dict(base_size=base, crop_size=crop, scales=scales, flip=True)
    # (1) the network inputs: every crop the fused kernel builds equals the torch chain (interpolate, pad, slice, pad,
    #     flip) to fp32 rounding — recorded through a network stand-in that just remembers what it is given
    seen = {}

    def recorder(tag):
        def f(x, label_set):
            seen.setdefault(tag, []).append(x.clone())
            return torch.zeros((x.shape[0], 1) + tuple(x.shape[2:]), device=x.device)
        return f
    MultiScaleEvaluator(recorder("glue"), fused=False, max_batch=1 << 30, **kw)(img, labels)
    MultiScaleEvaluator(recorder("fused"), fused=True, max_batch=1 << 30, **kw)(img, labels)
    a, b = torch.cat(seen["glue"], 0), torch.cat(seen["fused"], 0)
    assert a.shape == b.shape
    assert (a - b).abs().max().item() < 2e-6, (a - b).abs().max().item()
    # (2) end to end on the real network. Geometry / accumulation are exact (stand-in golden test below, 2e-5); here the
    #     network itself quantises its input to fp16, so inputs that differ by 1e-7 flip the rounding of a few hundred
    #     pixels per crop and the logits move like two evaluations of the fp16 pipeline do
    glue = MultiScaleEvaluator(net, fused=False, **kw)(img, labels)
    fused = MultiScaleEvaluator(net, fused=True, max_batch=16, **kw)(img, labels)
    assert fused.shape == glue.shape == (1, 3, h, w)
    assert rel_err(fused, glue) < 3e-3, rel_err(fused, glue)
    # the batched evaluation equals the sequential (one network input at a time) algorithm bit for bit
    seq = MultiScaleEvaluator(net, fused=True, max_batch=1, **kw)(img, labels)
    assert torch.equal(fused, seq)
    odd = MultiScaleEvaluator(net, fused=True, max_batch=3, **kw)(img, labels)
    assert torch.equal(fused, odd)
    # and the mask is what torch.max(scores, 1)[1] gives (test_lseg.py:397)
    assert torch.equal(MultiScaleEvaluator(net, **kw).predict(img, labels), torch.max(fused, 1)[1])


def test_fused_geometry_against_reference_golden():
    """tests/golden/ref_multiscale.npz holds the UNMODIFIED reference class's result around an exactly batch-invariant
    stand-in network (oracle/make_golden_eval.py): running the same stand-in on the GPU under the fused evaluator checks
    the kernels' geometry (window grid, flips, overlap average, resize back) against the reference itself."""
    import os
    from lseg_b200.evaluator import MultiScaleEvaluator
    from oracle.make_golden_eval import StandInNet
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_multiscale.npz"))
    n_cases = len([k for k in z.files if k.startswith("image")])
    assert n_cases >= 3
    for i in range(n_cases):
        h, w, base, crop, K, flip, seed = [int(v) for v in z[f"cfg{i}"]]
        standin = StandInNet(K, crop, seed)
        standin.Wc, standin.V, standin.pos = standin.Wc.cuda(), standin.V.cuda(), standin.pos.cuda()
        img = torch.from_numpy(z[f"image{i}"]).cuda()
        want = torch.from_numpy(z[f"scores{i}"])
        scales = [float(v) for v in z[f"scales{i}"]]
        got = MultiScaleEvaluator(standin, base_size=base, crop_size=crop, scales=scales, flip=bool(flip), fused=True)(
            img, ["c%d" % k for k in range(K)])
        assert got.shape == want.shape
        assert rel_err(got, want) < 2e-5, (i, rel_err(got, want))


def test_preprocess_kernel():
    """ToTensor + Normalize(.5,.5) + Resize([360,480]) of lseg_app.py:328-334, plus pad_image to the crop size."""
    from lseg_b200 import ops
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (333, 517, 3), generator=g, dtype=torch.uint8)
    x = img.permute(2, 0, 1).float().div(255.0)
    x = (x - 0.5) / 0.5
    ref = F.interpolate(x[None], size=(360, 480), mode="bilinear", align_corners=False)
    got = ops.preprocess(img.cuda(), (360, 480))
    assert got.shape == (1, 3, 360, 480)
    assert (got.cpu() - ref).abs().max().item() < 2e-6
    padded = ops.preprocess(img.cuda(), (360, 480), pad_to=(480, 480))
    assert torch.equal(padded[:, :, :360], got)
    assert (padded[:, :, 360:] == -1.0).all()
    # other statistics (ImageNet mean / std) and an up-scaling resize
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    x2 = (img.permute(2, 0, 1).float().div(255.0) - torch.tensor(mean)[:, None, None]) / torch.tensor(std)[:, None, None]
    ref2 = F.interpolate(x2[None], size=(700, 900), mode="bilinear", align_corners=False)
    got2 = ops.preprocess(img.cuda(), (700, 900), mean=mean, std=std)
    assert (got2.cpu() - ref2).abs().max().item() < 5e-6
