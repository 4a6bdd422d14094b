"""TEST INFRASTRUCTURE — CPU oracle: restatement of the reference LSeg forward path in plain torch.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (lang-seg_b200/) never does.

Parity status: the reference ships NO golden vectors or known-answer tests for this path
(SURVEY.md section 4), so the oracle is pinned against the reference itself: oracle/make_golden.py
imports /root/reference/modules/models/lseg_net.py unchanged (with stand-ins for the absent third-party
packages timm / clip, oracle/ref_standins.py), runs it on the seeded inputs of oracle/synth.py, checks
that this restatement reproduces it, and commits the outputs as fixtures under tests/golden/.
The timm-0.4.12 ViT block and the CLIP@04f4dc2 text tower are third-party code that is absent from
/root/reference; they are restated from their published algorithm and
cross-checked against transformers' CLIPTextModelWithProjection / ViTLayer in tests/test_oracle.py.

dtype pipeline = the reference's on CUDA: fp32 image trunk; fp16 text tower (clip.load(device='cuda')
converts Linear/MHA/text_projection weights to fp16, lseg_vit.py:224); fp16 pixel x text matmul
(lseg_net.py:194).
"""
import math

import torch
import torch.nn.functional as F

VIT_HOOKS = (5, 11, 17, 23)  # modules/models/lseg_net.py:119-123

# backbone -> (hooks, heads, per-level resampling after the 1x1 conv). lseg_net.py:119-123 (hooks),
# lseg_vit.py:221-238 + 408-532 (ViT-L/16: ConvT x4, ConvT x2, -, Conv 3x3 stride 2),
# lseg_vit.py:259-405 (ViT-B/32: ConvT x8, ConvT x4, ConvT x2, -). Patch size / width / depth are read off the weights.
# text_heads: heads of the CLIP text tower in use (ViT-B/32: 8; RN50x16: 12 — width 768, lseg_vit.py:243).
BACKBONES = {
    "clip_vitl16_384": dict(hooks=(5, 11, 17, 23), heads=16, resample=(4, 2, 0, -2), text_heads=8),
    "clipRN50x16_vitl16_384": dict(hooks=(5, 11, 17, 23), heads=16, resample=(4, 2, 0, -2), text_heads=12),
    "clip_vitb32_384": dict(hooks=(2, 5, 8, 11), heads=12, resample=(8, 4, 2, 0), text_heads=8),
}


# ------------------------------------------------------------------------------------------------
# image trunk
# ------------------------------------------------------------------------------------------------
def resize_pos_embed(posemb, gs_h, gs_w, start_index=1):
    """modules/models/lseg_vit.py:149-163 (bilinear, align_corners default False)."""
    posemb_tok, posemb_grid = posemb[:, :start_index], posemb[0, start_index:]
    gs_old = int(math.sqrt(len(posemb_grid)))
    posemb_grid = posemb_grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    posemb_grid = F.interpolate(posemb_grid, size=(gs_h, gs_w), mode="bilinear")
    posemb_grid = posemb_grid.permute(0, 2, 3, 1).reshape(1, gs_h * gs_w, -1)
    return torch.cat([posemb_tok, posemb_grid], dim=1)


def vit_attention(x, sd, prefix, num_heads=16):
    """timm 0.4.12 Attention.forward (same math as lseg_vit.py:26-39)."""
    B, N, C = x.shape
    qkv = F.linear(x, sd[prefix + "qkv.weight"], sd[prefix + "qkv.bias"])
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * ((C // num_heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def vit_block(x, sd, prefix, num_heads=16):
    """timm 0.4.12 Block.forward: x + Attn(LN1(x)); x + Mlp(LN2(x)); LN eps 1e-6, exact-erf GELU."""
    h = F.layer_norm(x, (x.shape[-1],), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-6)
    x = x + vit_attention(h, sd, prefix + "attn.", num_heads)
    h = F.layer_norm(x, (x.shape[-1],), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-6)
    h = F.linear(h, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"])
    return x + h


def vit_forward_flex(x, sd, hooks=VIT_HOOKS, depth=24, num_heads=16):
    """modules/models/lseg_vit.py:166-201 + the forward hooks of :421-426. Returns the four taps
    (outputs of blocks `hooks`, i.e. the un-normed residual stream). The final self.norm (:199) only
    feeds `glob`, which forward_vit discards (:108), so it is not computed."""
    p = "pretrained.model."
    b, c, h, w = x.shape
    patch = sd[p + "patch_embed.proj.weight"].shape[-1]  # model.patch_size (lseg_vit.py:295 / :526)
    pos_embed = resize_pos_embed(sd[p + "pos_embed"], h // patch, w // patch)
    x = F.conv2d(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)
    cls_tokens = sd[p + "cls_token"].expand(b, -1, -1)
    x = torch.cat((cls_tokens, x), dim=1)
    x = x + pos_embed
    taps = []
    for i in range(depth):
        x = vit_block(x, sd, f"{p}blocks.{i}.", num_heads)
        if i in hooks:
            taps.append(x)
        if i == max(hooks):
            break
    return taps


def project_readout(x, sd, prefix):
    """modules/models/lseg_vit.py:79-90: GELU(Linear(cat(tok, cls)))."""
    readout = x[:, 0].unsqueeze(1).expand_as(x[:, 1:])
    features = torch.cat((x[:, 1:], readout), -1)
    return F.gelu(F.linear(features, sd[prefix + "project.0.weight"], sd[prefix + "project.0.bias"]))


def forward_vit(x, sd, taps_out=None, backbone="clip_vitl16_384"):
    """modules/models/lseg_vit.py:104-146 with the act_postprocess stacks of :442-522 (ViT-L/16) / :309-398 (ViT-B/32)."""
    cfg = BACKBONES[backbone]
    b, c, h, w = x.shape
    patch = sd["pretrained.model.patch_embed.proj.weight"].shape[-1]
    taps = vit_forward_flex(x, sd, cfg["hooks"], max(cfg["hooks"]) + 1, cfg["heads"])
    if taps_out is not None:
        taps_out.extend(taps)
    layers = []
    for k, tap in enumerate(taps):
        q = f"pretrained.act_postprocess{k + 1}."
        y = project_readout(tap, sd, q + "0.").transpose(1, 2)
        y = y.unflatten(2, (h // patch, w // patch))
        y = F.conv2d(y, sd[q + "3.weight"], sd[q + "3.bias"])
        r = cfg["resample"][k]
        if r > 0:
            y = F.conv_transpose2d(y, sd[q + "4.weight"], sd[q + "4.bias"], stride=r)
        elif r == -2:
            y = F.conv2d(y, sd[q + "4.weight"], sd[q + "4.bias"], stride=2, padding=1)
        layers.append(y)
    return layers


# ------------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------------
def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd[prefix + "weight"],
                        sd[prefix + "bias"], False, 0.0, 1e-5)


def residual_conv_unit(x, sd, prefix):
    """modules/models/lseg_blocks.py:265-288 (bn=True, activation = ReLU(inplace=False))."""
    out = F.relu(x)
    out = _bn(F.conv2d(out, sd[prefix + "conv1.weight"], None, padding=1), sd, prefix + "bn1.")
    out = F.relu(out)
    out = _bn(F.conv2d(out, sd[prefix + "conv2.weight"], None, padding=1), sd, prefix + "bn2.")
    return out + x


def feature_fusion_block(sd, prefix, *xs):
    """modules/models/lseg_blocks.py:337-358."""
    output = xs[0]
    if len(xs) == 2:
        output = output + residual_conv_unit(xs[1], sd, prefix + "resConfUnit1.")
    output = residual_conv_unit(output, sd, prefix + "resConfUnit2.")
    output = F.interpolate(output, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(output, sd[prefix + "out_conv.weight"], sd[prefix + "out_conv.bias"])


def decoder(layers, sd):
    """modules/models/lseg_net.py:171-179."""
    rn = [F.conv2d(layers[k], sd[f"scratch.layer{k + 1}_rn.weight"], None, padding=1) for k in range(4)]
    path_4 = feature_fusion_block(sd, "scratch.refinenet4.", rn[3])
    path_3 = feature_fusion_block(sd, "scratch.refinenet3.", path_4, rn[2])
    path_2 = feature_fusion_block(sd, "scratch.refinenet2.", path_3, rn[1])
    path_1 = feature_fusion_block(sd, "scratch.refinenet1.", path_2, rn[0])
    return path_1


# ------------------------------------------------------------------------------------------------
# CLIP text tower (fp16, as loaded on CUDA)
# ------------------------------------------------------------------------------------------------
def _ln_fp32(x, w, b, eps=1e-5):
    """CLIP's LayerNorm subclass: computed in fp32, cast back to the input dtype."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps).to(x.dtype)


def clip_text_weights_fp16(sd, dtype=torch.float16):
    """clip.model.convert_weights: Linear / MultiheadAttention / text_projection -> fp16; LayerNorm,
    embeddings stay fp32. dtype=float32 gives the un-converted tower (tests)."""
    c = "clip_pretrained."
    out = {}
    for k, v in sd.items():
        if not k.startswith(c) or k.startswith(c + "visual."):
            continue
        name = k[len(c):]
        if ("ln_" in name) or name in ("positional_embedding", "token_embedding.weight", "logit_scale"):
            out[name] = v.float()
        else:
            out[name] = v.to(dtype)
    return out


def clip_encode_text(text, tw, heads=8, layers=12, dtype=torch.float16, layer_io=None, op_trace=None):
    """CLIP.encode_text. text int64 [K,77] -> fp16 [K,512].
    layer_io: optional list that receives the residual stream [K,77,512] before block 0 and after every block
    (teacher-forced per-block parity tests). op_trace: optional dict {block index: {}} that receives, for those blocks,
    the output of every op of the block ([K,77,*] layout): ln1, qkv, attn (heads merged), x1, ln2, gelu, x2."""
    x = tw["token_embedding.weight"][text].to(dtype)
    x = x + tw["positional_embedding"].to(dtype)
    K, L, Wd = x.shape
    mask = torch.full((L, L), float("-inf")).triu_(1).to(dtype)
    x = x.permute(1, 0, 2)  # LND
    if layer_io is not None:
        layer_io.append(x.permute(1, 0, 2).clone())
    for i in range(layers):
        b = f"transformer.resblocks.{i}."
        tr = op_trace.get(i) if op_trace is not None else None
        h = _ln_fp32(x, tw[b + "ln_1.weight"], tw[b + "ln_1.bias"])
        # nn.MultiheadAttention (torch 1.9 multi_head_attention_forward)
        qkv = F.linear(h, tw[b + "attn.in_proj_weight"], tw[b + "attn.in_proj_bias"])
        if tr is not None:
            tr["ln1"], tr["qkv"] = h.permute(1, 0, 2).clone(), qkv.permute(1, 0, 2).clone()
        q, k, v = qkv.chunk(3, dim=-1)
        hd = Wd // heads
        q = q * (float(hd) ** -0.5)
        q = q.contiguous().view(L, K * heads, hd).transpose(0, 1)
        k = k.contiguous().view(L, K * heads, hd).transpose(0, 1)
        v = v.contiguous().view(L, K * heads, hd).transpose(0, 1)
        attn = torch.bmm(q, k.transpose(1, 2)) + mask
        attn = F.softmax(attn, dim=-1)
        o = torch.bmm(attn, v).transpose(0, 1).contiguous().view(L, K, Wd)
        if tr is not None:
            tr["attn"] = o.permute(1, 0, 2).clone()
        o = F.linear(o, tw[b + "attn.out_proj.weight"], tw[b + "attn.out_proj.bias"])
        x = x + o
        h = _ln_fp32(x, tw[b + "ln_2.weight"], tw[b + "ln_2.bias"])
        if tr is not None:
            tr["x1"], tr["ln2"] = x.permute(1, 0, 2).clone(), h.permute(1, 0, 2).clone()
        h = F.linear(h, tw[b + "mlp.c_fc.weight"], tw[b + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)  # QuickGELU
        if tr is not None:
            tr["gelu"] = h.permute(1, 0, 2).clone()
        h = F.linear(h, tw[b + "mlp.c_proj.weight"], tw[b + "mlp.c_proj.bias"])
        x = x + h
        if tr is not None:
            tr["x2"] = x.permute(1, 0, 2).clone()
        if layer_io is not None:
            layer_io.append(x.permute(1, 0, 2).clone())
    x = x.permute(1, 0, 2)
    x = _ln_fp32(x, tw["ln_final.weight"], tw["ln_final.bias"]).to(dtype)
    x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ tw["text_projection"]
    return x


# ------------------------------------------------------------------------------------------------
# head
# ------------------------------------------------------------------------------------------------
LOGIT_SCALE = torch.tensor(math.log(1 / 0.07)).exp()  # modules/models/lseg_net.py:141


def correlation_head(path_1, text_features, sd):
    """modules/models/lseg_net.py:185-196: head1, L2 norms, fp16 scale-then-matmul, NCHW fp32 view."""
    image_features = F.conv2d(path_1, sd["scratch.head1.weight"], sd["scratch.head1.bias"])
    imshape = image_features.shape
    image_features = image_features.permute(0, 2, 3, 1).reshape(-1, imshape[1])
    image_features = image_features / image_features.norm(dim=-1, keepdim=True)
    text_features = text_features / text_features.norm(dim=-1, keepdim=True)
    logits_per_image = LOGIT_SCALE * image_features.half() @ text_features.t()
    return logits_per_image.float().view(imshape[0], imshape[2], imshape[3], -1).permute(0, 3, 1, 2)


def head_block(out, sd, arch_option, block_depth, activation):
    """scratch.head_block applied as lseg_net.py:198-201 does: `block_depth - 1` times with the activation, once without.
    depthwise_conv (lseg_net.py:29-42) folds the classes into the batch and runs ONE Conv2d(1,1,3,padding=1) over every
    class plane; bottleneck_block (:62-79) adds the per-pixel max over the classes of its input; depthwise_block (:45-60)
    does not."""
    w = sd["scratch.head_block.depthwise.depthwise.weight"]
    b = sd["scratch.head_block.depthwise.depthwise.bias"]
    act = {"relu": F.relu, "lrelu": lambda t: F.leaky_relu(t, 0.01), "tanh": torch.tanh}[activation]

    def block(x, use_act):
        B, C, H, W = x.shape
        y = F.conv2d(x.reshape(-1, 1, H, W), w, b, padding=1).view(-1, C, H, W)
        if arch_option == 1:
            y = y + x.max(dim=1, keepdim=True)[0]
        return act(y) if use_act else y

    for _ in range(block_depth - 1):
        out = block(out, True)
    return block(out, False)


def output_conv(out):
    """scratch.output_conv = Interpolate(x2, bilinear, align_corners=True) (lseg_net.py:203,219-221)."""
    return F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)


@torch.no_grad()
def lseg_forward(x, tokens, sd, text_weights=None, return_stages=False, arch_option=0, block_depth=0,
                 activation="lrelu", backbone="clip_vitl16_384"):
    """LSeg.forward (modules/models/lseg_net.py:160-205). x fp32 [B,3,H,W], tokens int64 [K,77]."""
    if x.shape[2] % 32 or x.shape[3] % 32:
        raise ValueError("H and W must be multiples of 32 (even token grid)")
    tw = text_weights if text_weights is not None else clip_text_weights_fp16(sd)
    taps = []
    layers = forward_vit(x, sd, taps, backbone)
    path_1 = decoder(layers, sd)
    text_features = clip_encode_text(tokens, tw, heads=BACKBONES[backbone]["text_heads"])
    low = correlation_head(path_1, text_features, sd)
    if arch_option in (1, 2):
        low = head_block(low, sd, arch_option, block_depth, activation)
    out = output_conv(low)
    if return_stages:
        tf = text_features / text_features.norm(dim=-1, keepdim=True)
        return out, {"taps": taps, "layers": layers, "path_1": path_1, "text_features": tf, "logits_lr": low}
    return out


# ------------------------------------------------------------------------------------------------
# ResNet-101 trunk of the zero-shot model LSegRNNetZS (modules/models/lseg_net_zs.py:240-339; backbone
# "clip_resnet101": torchvision resnet101 split by _make_resnet_backbone, lseg_blocks_zs.py:109-119 /
# lseg_vit_zs.py:742-760). Third-party arithmetic restated: torchvision 0.10 (torch 1.9.1) ResNet v1.5 — Bottleneck =
# 1x1 -> BN -> ReLU -> 3x3 (stride on THIS conv) -> BN -> ReLU -> 1x1 -> BN, + identity (1x1 stride-s conv + BN when the
# shape changes), ReLU; stem 7x7 s2 p3 -> BN -> ReLU -> maxpool 3x3 s2 p1; BatchNorm eval, eps 1e-5.
# ------------------------------------------------------------------------------------------------
RESNET101_LAYERS = (3, 4, 23, 3)


def _rn_prefix(layer, i):
    """state-dict prefix of block i of layer 1..4: layer1 is nn.Sequential(conv1, bn1, relu, maxpool, resnet.layer1)."""
    return f"pretrained.layer1.4.{i}." if layer == 1 else f"pretrained.layer{layer}.{i}."


def resnet_bottleneck(x, sd, prefix, stride):
    out = F.relu(_bn(F.conv2d(x, sd[prefix + "conv1.weight"]), sd, prefix + "bn1."))
    out = F.relu(_bn(F.conv2d(out, sd[prefix + "conv2.weight"], stride=stride, padding=1), sd, prefix + "bn2."))
    out = _bn(F.conv2d(out, sd[prefix + "conv3.weight"]), sd, prefix + "bn3.")
    if prefix + "downsample.0.weight" in sd:
        x = _bn(F.conv2d(x, sd[prefix + "downsample.0.weight"], stride=stride), sd, prefix + "downsample.1.")
    return F.relu(out + x)


def resnet101_layers(x, sd):
    """pretrained.layer1..4 of lseg_net_zs.py:307-310 -> four NCHW fp32 maps (256/512/1024/2048 channels, strides 4..32)."""
    x = F.relu(_bn(F.conv2d(x, sd["pretrained.layer1.0.weight"], stride=2, padding=3), sd, "pretrained.layer1.1."))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for layer, n in enumerate(RESNET101_LAYERS, start=1):
        for i in range(n):
            x = resnet_bottleneck(x, sd, _rn_prefix(layer, i), 2 if (layer > 1 and i == 0) else 1)
        outs.append(x)
    return outs


@torch.no_grad()
def lseg_forward_rn_zs(x, class_info, label_tokens, sd, text_weights=None, return_stages=False):
    """LSegRN.forward (modules/models/lseg_net_zs.py:300-338): ResNet-101 layers -> the same scratch decoder and the
    zero-shot head (one ['others', name] pair per image)."""
    tw = text_weights if text_weights is not None else clip_text_weights_fp16(sd)
    layers = resnet101_layers(x, sd)
    path_1 = decoder(layers, sd)
    outs = []
    for i in range(x.shape[0]):
        tf = clip_encode_text(label_tokens[int(class_info[i])], tw)
        outs.append(correlation_head(path_1[i:i + 1], tf, sd))
    out = output_conv(torch.cat(outs, dim=0))
    if return_stages:
        return out, {"layers": layers, "path_1": path_1}
    return out


@torch.no_grad()
def lseg_forward_zs(x, class_info, label_tokens, sd, text_weights=None):
    """Zero-shot LSeg.forward (modules/models/lseg_net_zs.py:177-214): per-image ['others', name] pair.
    label_tokens: list of int64 [2,77] tensors (self.texts, :169-175); class_info int64 [B]."""
    tw = text_weights if text_weights is not None else clip_text_weights_fp16(sd)
    layers = forward_vit(x, sd)
    path_1 = decoder(layers, sd)
    outs = []
    for i in range(x.shape[0]):
        tf = clip_encode_text(label_tokens[int(class_info[i])], tw)
        outs.append(correlation_head(path_1[i:i + 1], tf, sd))
    return output_conv(torch.cat(outs, dim=0))
