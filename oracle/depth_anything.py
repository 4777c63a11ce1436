"""ORACLE (test infrastructure only - never imported by nunif_b200/): CPU/torch restatement of the Depth-Anything-V2
network the reference loads through torch.hub (iw3/depth_anything_model.py:223-225,
"nagadomi/Depth-Anything_iw3:main", DepthAnything(encoder="v2_vits")) - THIRD-PARTY code that is not in
/root/reference and cannot be fetched here.

PARITY UNPINNED against the reference's own copy: this file restates the published architecture
(DINOv2 ViT-S/14 encoder, facebookresearch/dinov2 vision_transformer.py; DPT head of
DepthAnything/Depth-Anything-V2 depth_anything_v2/dpt.py) with the upstream state_dict key names
(`pretrained.*`, `depth_head.*`).  It is cross-checked numerically against an independent public implementation of
the same architecture that IS in this image (transformers.DepthAnythingForDepthEstimation,
tests/test_oracle_golden.py::test_depth_anything_oracle_matches_transformers) and anchored on the reference's call
sites: batch_preprocess -> model(x) under fp16 autocast -> unsqueeze(1).float() (depth_anything_model.py:113-182).

Functional style (state_dict in, tensors out) so the same code runs in fp32 on the CPU or under CUDA autocast in the
GPU tests (the reference's numerics, nunif/device.py:58-71).
"""
import math
import torch
import torch.nn.functional as F

PATCH = 14
ENCODERS = {
    # name: (dim, depth, heads, intermediate layers, head features, head out_channels)
    "vits": (384, 12, 6, (2, 5, 8, 11), 64, (48, 96, 192, 384)),
    "vitb": (768, 12, 12, (2, 5, 8, 11), 128, (96, 192, 384, 768)),
    "vitl": (1024, 24, 16, (4, 11, 17, 23), 256, (256, 512, 1024, 1024)),
}


def interpolate_pos_encoding(pos_embed, H, W, offset=0.1):
    """dinov2 vision_transformer.py interpolate_pos_encoding (interpolate_offset=0.1, antialias off): the learned
    (1, 1+M*M, dim) table -> (1, 1+(H/14)*(W/14), dim) for an H x W input."""
    N = pos_embed.shape[1] - 1
    ph, pw = H // PATCH, W // PATCH
    if ph * pw == N and H == W:
        return pos_embed
    dim = pos_embed.shape[-1]
    M = int(math.sqrt(N))
    assert M * M == N
    sy, sx = float(ph + offset) / M, float(pw + offset) / M
    pp = pos_embed[:, 1:].float().reshape(1, M, M, dim).permute(0, 3, 1, 2)
    pp = F.interpolate(pp, scale_factor=(sy, sx), mode="bicubic", antialias=False)
    assert pp.shape[-2] == ph and pp.shape[-1] == pw
    pp = pp.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat([pos_embed[:, :1].float(), pp], dim=1).to(pos_embed.dtype)


def vit_intermediate(sd, x, encoder="vits"):
    """get_intermediate_layers(x, idx, return_class_token=True) with the final norm applied: list of
    (patch tokens B,N,dim ; class token B,dim)."""
    dim, depth, heads, idx, _, _ = ENCODERS[encoder]
    B, _, H, W = x.shape
    t = F.conv2d(x, sd["pretrained.patch_embed.proj.weight"], sd["pretrained.patch_embed.proj.bias"], stride=PATCH)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd["pretrained.cls_token"].expand(B, -1, -1).to(t.dtype), t], dim=1)
    t = t + interpolate_pos_encoding(sd["pretrained.pos_embed"], H, W).to(t.dtype)
    d = dim // heads
    outs = []
    for i in range(depth):
        p = f"pretrained.blocks.{i}."
        h = F.layer_norm(t, (dim,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-6)
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        qkv = qkv.reshape(B, -1, 3, heads, d).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * (d ** -0.5), qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        h = (a @ v).transpose(1, 2).reshape(B, -1, dim)
        h = F.linear(h, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        t = t + h * sd[p + "ls1.gamma"]
        h = F.layer_norm(t, (dim,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps=1e-6)
        h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                     sd[p + "mlp.fc2.bias"])
        t = t + h * sd[p + "ls2.gamma"]
        if i in idx:
            outs.append(t)
    outs = [F.layer_norm(o, (dim,), sd["pretrained.norm.weight"], sd["pretrained.norm.bias"], eps=1e-6) for o in outs]
    return [(o[:, 1:], o[:, 0]) for o in outs]


def _rcu(sd, p, x):
    """ResidualConvUnit (dpt blocks.py): relu - conv3x3 - relu - conv3x3 + x."""
    out = F.conv2d(F.relu(x), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    out = F.conv2d(F.relu(out), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    return out + x


def _fusion(sd, p, x0, x1=None, size=None):
    """FeatureFusionBlock.forward."""
    out = x0
    if x1 is not None:
        out = out + _rcu(sd, p + "resConfUnit1.", x1)
    out = _rcu(sd, p + "resConfUnit2.", out)
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def dpt_head(sd, feats, ph, pw, encoder="vits"):
    """DPTHead.forward (use_clstoken=False) -> B,1,ph*14,pw*14 (after the model-level ReLU)."""
    dim = ENCODERS[encoder][0]
    layers = []
    for i, (x, _cls) in enumerate(feats):
        B = x.shape[0]
        x = x.permute(0, 2, 1).reshape(B, dim, ph, pw)
        x = F.conv2d(x, sd[f"depth_head.projects.{i}.weight"], sd[f"depth_head.projects.{i}.bias"])
        if i == 0:
            x = F.conv_transpose2d(x, sd["depth_head.resize_layers.0.weight"], sd["depth_head.resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, sd["depth_head.resize_layers.1.weight"], sd["depth_head.resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = F.conv2d(x, sd["depth_head.resize_layers.3.weight"], sd["depth_head.resize_layers.3.bias"], stride=2, padding=1)
        layers.append(F.conv2d(x, sd[f"depth_head.scratch.layer{i + 1}_rn.weight"], None, padding=1))
    l1, l2, l3, l4 = layers
    s = "depth_head.scratch."
    p4 = _fusion(sd, s + "refinenet4.", l4, None, size=l3.shape[2:])
    p3 = _fusion(sd, s + "refinenet3.", p4, l3, size=l2.shape[2:])
    p2 = _fusion(sd, s + "refinenet2.", p3, l2, size=l1.shape[2:])
    p1 = _fusion(sd, s + "refinenet1.", p2, l1)
    out = F.conv2d(p1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)
    out = F.interpolate(out, (ph * PATCH, pw * PATCH), mode="bilinear", align_corners=True)
    out = F.relu(F.conv2d(out, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1))
    out = F.relu(F.conv2d(out, sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"]))
    return F.relu(out)


def depth_anything_forward(sd, x, encoder="vits"):
    """DepthAnythingV2.forward: x B,3,H,W (ImageNet-normalised, H and W multiples of 14) -> depth B,H,W."""
    ph, pw = x.shape[-2] // PATCH, x.shape[-1] // PATCH
    feats = vit_intermediate(sd, x, encoder)
    return dpt_head(sd, feats, ph, pw, encoder).squeeze(1)
