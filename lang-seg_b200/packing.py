"""State-dict -> packed device buffers for the sm_100a engine (host-side, one-off per weight load).

Input: the reference's state dict (key names of SURVEY.md Appendix C, fp32). Output: the
`lseg_weights` descriptor of include/lseg_b200.h plus the tensors that back its pointers.

Layouts
  * every GEMM weight is fp16 [N, K] row-major (K contiguous = "K-major" UMMA operand), rows padded
    with zeros to a multiple of 128 so a TMA box never exceeds the tensor;
  * 3x3 convs are tap-major: [N, (ky, kx, c)]  (the implicit-GEMM K loop walks taps, then channels);
  * ConvTranspose2d with kernel == stride becomes a GEMM with N = (i, j, co) and a depth-to-space
    store; its bias is expanded to [s*s*cout];
  * ProjectReadout's Linear(2048->1024) is split into the token half W[:, :1024] and the cls half
    W[:, 1024:] (+bias) — exact up to summation order;
  * eval-mode BatchNorm is folded to per-channel (scale, shift) applied in the GEMM epilogue in fp32.
"""
import math

import numpy as np
import torch

from . import _lib

VIT_DEPTH = _lib.VIT_DEPTH
TEXT_DEPTH = _lib.TEXT_DEPTH
HOOKS = (5, 11, 17, 23)  # modules/models/lseg_net.py:119-123


def _pad_rows(w, mult=128):
    n = w.shape[0]
    npad = (n + mult - 1) // mult * mult
    if npad == n:
        return w.contiguous()
    out = torch.zeros((npad,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
    out[:n] = w
    return out


def conv3x3_to_gemm(w):
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin] tap-major (ky, kx, c)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def deconv_to_gemm(w):
    """ConvTranspose2d weight [Cin, Cout, s, s] (kernel == stride) -> [(i, j, co), ci]."""
    cin, cout, s, _ = w.shape
    return w.permute(2, 3, 1, 0).reshape(s * s * cout, cin)


def fold_bn(sd, prefix, eps=1e-5):
    """BatchNorm2d eval -> y = x*scale + shift (modules/models/lseg_blocks.py:265-288, eps 1e-5)."""
    g, b = sd[prefix + "weight"].float(), sd[prefix + "bias"].float()
    m, v = sd[prefix + "running_mean"].float(), sd[prefix + "running_var"].float()
    scale = g / torch.sqrt(v + eps)
    return scale, b - m * scale


def reference_logit_scale():
    """exp(log(1/0.07)) evaluated in fp32 like `nn.Parameter(torch.ones([]) * np.log(1 / 0.07)).exp()`
    (modules/models/lseg_net.py:141)."""
    return float((torch.ones([]) * np.log(1 / 0.07)).exp())


class PackedWeights:
    """Owns the packed tensors and the ctypes descriptor that points into them."""

    def __init__(self, state_dict, device, arch_option=0, block_depth=0, activation="lrelu"):
        self.device = torch.device(device)
        self._keep = []
        self.desc = _lib.Weights()
        self._pack(state_dict)
        d = self.desc
        d.arch_option, d.block_depth = int(arch_option or 0), int(block_depth or 0)
        d.head_act = _lib.HEAD_ACT[activation] if d.arch_option else 0
        if d.arch_option in (1, 2):  # scratch.head_block (lseg_net.py:148-154): one shared 3x3 kernel + bias
            w = state_dict["scratch.head_block.depthwise.depthwise.weight"].detach().float().reshape(9).tolist()
            for i in range(9):
                d.head_block_w[i] = w[i]
            d.head_block_b = float(state_dict["scratch.head_block.depthwise.depthwise.bias"].detach().float().reshape(-1)[0])

    # -- helpers -------------------------------------------------------------------------------
    def _f32(self, t):
        t = t.detach().to(self.device, torch.float32).contiguous()
        self._keep.append(t)
        return t.data_ptr()

    def _f16_rows(self, t):
        t = _pad_rows(t.detach().to(self.device, torch.float32).to(torch.float16).contiguous())
        self._keep.append(t)
        return t

    def _linear(self, slot, w2d, bias, round_bias_f16=False):
        wt = self._f16_rows(w2d)
        slot.w = wt.data_ptr()
        slot.out, slot.in_ = int(w2d.shape[0]), int(w2d.shape[1])
        slot.rows = int(wt.shape[0])
        if bias is not None:
            b = bias.detach().float()
            if round_bias_f16:  # CLIP's Linear biases are fp16 tensors after convert_weights
                b = b.half().float()
            slot.b = self._f32(b)
        else:
            slot.b = None

    # -- packing -------------------------------------------------------------------------------
    def _pack(self, sd):
        d = self.desc
        p = "pretrained.model."
        self._linear(d.patch, sd[p + "patch_embed.proj.weight"].reshape(1024, 768), sd[p + "patch_embed.proj.bias"])
        d.cls_token = self._f32(sd[p + "cls_token"].reshape(1024))
        pos = sd[p + "pos_embed"].reshape(-1, 1024)
        d.pos_embed = self._f32(pos)
        d.pos_grid = int(round(math.sqrt(pos.shape[0] - 1)))
        for i in range(VIT_DEPTH):
            b = f"{p}blocks.{i}."
            blk = d.blocks[i]
            blk.ln1_g, blk.ln1_b = self._f32(sd[b + "norm1.weight"]), self._f32(sd[b + "norm1.bias"])
            blk.ln2_g, blk.ln2_b = self._f32(sd[b + "norm2.weight"]), self._f32(sd[b + "norm2.bias"])
            self._linear(blk.qkv, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"])
            self._linear(blk.proj, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
            self._linear(blk.fc1, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"])
            self._linear(blk.fc2, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
        for k in range(4):
            d.hooks[k] = HOOKS[k]
            q = f"pretrained.act_postprocess{k + 1}."
            w = sd[q + "0.project.0.weight"]
            self._linear(d.readout_tok[k], w[:, :1024], None)
            self._linear(d.readout_cls[k], w[:, 1024:], sd[q + "0.project.0.bias"])
            c = sd[q + "3.weight"]
            self._linear(d.post_conv1x1[k], c.reshape(c.shape[0], c.shape[1]), sd[q + "3.bias"])
        w = sd["pretrained.act_postprocess1.4.weight"]
        self._linear(d.post1_deconv, deconv_to_gemm(w), sd["pretrained.act_postprocess1.4.bias"].repeat(16))
        w = sd["pretrained.act_postprocess2.4.weight"]
        self._linear(d.post2_deconv, deconv_to_gemm(w), sd["pretrained.act_postprocess2.4.bias"].repeat(4))
        self._linear(d.post4_conv, conv3x3_to_gemm(sd["pretrained.act_postprocess4.4.weight"]),
                     sd["pretrained.act_postprocess4.4.bias"])
        for k in range(4):
            self._linear(d.layer_rn[k], conv3x3_to_gemm(sd[f"scratch.layer{k + 1}_rn.weight"]), None)
            q = f"scratch.refinenet{k + 1}."
            for slot, unit in ((d.rcu1[k], "resConfUnit1."), (d.rcu2[k], "resConfUnit2.")):
                self._linear(slot.conv1, conv3x3_to_gemm(sd[q + unit + "conv1.weight"]), None)
                self._linear(slot.conv2, conv3x3_to_gemm(sd[q + unit + "conv2.weight"]), None)
                s1, t1 = fold_bn(sd, q + unit + "bn1.")
                s2, t2 = fold_bn(sd, q + unit + "bn2.")
                slot.bn1_scale, slot.bn1_shift = self._f32(s1), self._f32(t1)
                slot.bn2_scale, slot.bn2_shift = self._f32(s2), self._f32(t2)
            oc = sd[q + "out_conv.weight"]
            self._linear(d.out_conv[k], oc.reshape(oc.shape[0], oc.shape[1]), sd[q + "out_conv.bias"])
        h1 = sd["scratch.head1.weight"]
        self._linear(d.head1, h1.reshape(h1.shape[0], h1.shape[1]), sd["scratch.head1.bias"])
        d.logit_scale = reference_logit_scale()
        # CLIP text tower: Linear / MHA weights are fp16 in the reference (clip.load on cuda)
        c = "clip_pretrained."
        d.tok_emb = self._f32(sd[c + "token_embedding.weight"])
        d.text_pos = self._f32(sd[c + "positional_embedding"])
        for i in range(TEXT_DEPTH):
            b = f"{c}transformer.resblocks.{i}."
            blk = d.text_blocks[i]
            blk.ln1_g, blk.ln1_b = self._f32(sd[b + "ln_1.weight"]), self._f32(sd[b + "ln_1.bias"])
            blk.ln2_g, blk.ln2_b = self._f32(sd[b + "ln_2.weight"]), self._f32(sd[b + "ln_2.bias"])
            self._linear(blk.in_proj, sd[b + "attn.in_proj_weight"], sd[b + "attn.in_proj_bias"], True)
            self._linear(blk.out_proj, sd[b + "attn.out_proj.weight"], sd[b + "attn.out_proj.bias"], True)
            self._linear(blk.c_fc, sd[b + "mlp.c_fc.weight"], sd[b + "mlp.c_fc.bias"], True)
            self._linear(blk.c_proj, sd[b + "mlp.c_proj.weight"], sd[b + "mlp.c_proj.bias"], True)
        d.lnf_g, d.lnf_b = self._f32(sd[c + "ln_final.weight"]), self._f32(sd[c + "ln_final.bias"])
        self._linear(d.text_proj, sd[c + "text_projection"].t(), None)
