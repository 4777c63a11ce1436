"""ORACLE (test infrastructure only - never imported by nunif_b200/): CPU/torch restatement of the ZoeD_N metric depth
network the reference loads through torch.hub (iw3/zoedepth_model.py:151-157, "nagadomi/ZoeDepth_iw3:main", ZoeD_N,
config_mode="infer") - THIRD-PARTY code (isl-org/ZoeDepth + isl-org/MiDaS DPT_BEiT_L_384) that is not in
/root/reference and cannot be fetched here.

PARITY UNPINNED against the reference's own copy: this file restates the published architecture
  * BEiT-L/16 encoder with per-layer relative position bias, resampled for non-square token grids
    (MiDaS v3.1 backbones/beit.py `_get_rel_pos_bias`, hooks after blocks 5, 11, 17, 23),
  * DPT reassemble ("project" readout) + RefineNet fusion + output_conv (MiDaS dpt_depth.py / blocks.py),
  * the metric bins head: seed regressor, 4 unnormed attractor layers, conditional log-binomial
    (ZoeDepth zoedepth_v1.py, layers/attractor.py, layers/dist_layers.py, layers/localbins_layers.py)
with the upstream checkpoint key names of ZoeD_M12_N.pt (`core.core.pretrained.*`, `core.core.scratch.*`, `conv2`,
`seed_bin_regressor`, `seed_projector`, `projectors`, `attractors`, `conditional_log_binomial`; names restated from the
published conversion table between the upstream checkpoint and transformers).  It is cross-checked numerically against
an independent public implementation of the same architecture that IS in this image
(transformers.ZoeDepthForDepthEstimation, verified by its authors against the upstream weights;
tests/test_oracle_golden.py::test_zoedepth_oracle_matches_transformers) and anchored on the reference's call sites:
batch_preprocess -> model(x)['metric_depth'] under fp16 autocast -> nan_to_num (zoedepth_model.py:23-27, 89-148).

Two upstream quirks are kept because the released weights were trained with them (both are also in transformers):
  * the attractor layers call `inv_attractor(dx)` with its DEFAULT alpha = 300, gamma = 2 (not the configured 1000),
  * the relative-position sub-table is reshaped as (old_width, old_height) before the bilinear resample.

Functional style (state_dict in, tensors out) so the same code runs in fp32 on the CPU or under CUDA autocast in the GPU
tests (the reference's numerics, nunif/device.py:58-71).
"""
import math
import torch
import torch.nn.functional as F

PATCH = 16
# ZoeD_N: BEiT-L/16 (384 training grid = 24 x 24 tokens), DPT_BEiT_L_384 head
ZOED_N = dict(dim=1024, depth=24, heads=16, hooks=(5, 11, 17, 23), oc=(256, 512, 1024, 1024), feat=256, old_grid=24)
# reduced configuration for fast CPU / GPU tests (same code path, every width a multiple of 64)
ZOED_MINI = dict(dim=256, depth=4, heads=4, hooks=(0, 1, 2, 3), oc=(64, 128, 256, 256), feat=128, old_grid=6)
N_BINS = 64
N_ATTRACTORS = (16, 8, 4, 1)
BIN_DIM = 128
MIN_TEMP, MAX_TEMP = 0.0212, 50.0
BB = "core.core.pretrained.model."
PP = "core.core.pretrained."
SC = "core.core.scratch."


def relative_position_index(wh, ww):
    """timm/MiDaS beit.py gen_relative_position_index for a wh x ww token grid + the 3 class-token entries."""
    nrd = (2 * wh - 1) * (2 * ww - 1) + 3
    coords = torch.stack(torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += wh - 1
    rel[:, :, 1] += ww - 1
    rel[:, :, 0] *= 2 * ww - 1
    idx = torch.zeros((wh * ww + 1,) * 2, dtype=rel.dtype)
    idx[1:, 1:] = rel.sum(-1)
    idx[0, 0:] = nrd - 3
    idx[0:, 0] = nrd - 2
    idx[0, 0] = nrd - 1
    return idx


def relative_position_bias(table, old_grid, wh, ww):
    """MiDaS backbones/beit.py _get_rel_pos_bias: learned table ((2g-1)^2 + 3, heads) -> bias (heads, 1+wh*ww, 1+wh*ww)."""
    oh = ow = 2 * old_grid - 1
    nh, nw = 2 * wh - 1, 2 * ww - 1
    n_old = oh * ow + 3
    assert table.shape[0] == n_old
    sub = table[:n_old - 3].float().reshape(1, ow, oh, -1).permute(0, 3, 1, 2)
    new = F.interpolate(sub, size=(nh, nw), mode="bilinear")
    new = new.permute(0, 2, 3, 1).reshape(nh * nw, -1)
    full = torch.cat([new, table[n_old - 3:].float()])
    idx = relative_position_index(wh, ww).to(table.device)
    n = wh * ww + 1
    return full[idx.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()


def beit_features(sd, x, cfg=ZOED_N):
    """Hidden states (B, 1+P, DIM) after blocks HOOKS; no final norm (MiDaS forward hooks on the raw block outputs)."""
    B, _, H, W = x.shape
    ph, pw = H // PATCH, W // PATCH
    DIM, DEPTH, HEADS, HOOKS, old_grid = cfg["dim"], cfg["depth"], cfg["heads"], cfg["hooks"], cfg["old_grid"]
    t = F.conv2d(x, sd[BB + "patch_embed.proj.weight"], sd[BB + "patch_embed.proj.bias"], stride=PATCH)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd[BB + "cls_token"].expand(B, -1, -1).to(t.dtype), t], dim=1)
    d = DIM // HEADS
    outs = []
    for i in range(DEPTH):
        p = f"{BB}blocks.{i}."
        h = F.layer_norm(t, (DIM,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv_bias = torch.cat([sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.v_bias"]), sd[p + "attn.v_bias"]])
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], qkv_bias).reshape(B, -1, 3, HEADS, d).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) / math.sqrt(d)
        attn = attn + relative_position_bias(sd[p + "attn.relative_position_bias_table"], old_grid, ph, pw).unsqueeze(0)
        attn = attn.softmax(dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(B, -1, DIM)
        o = F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        t = t + sd[p + "gamma_1"] * o
        h = F.layer_norm(t, (DIM,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        t = t + sd[p + "gamma_2"] * h
        if i in HOOKS:
            outs.append(t)
    return outs


def _rcu(sd, p, x):
    y = F.conv2d(F.relu(x), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    y = F.conv2d(F.relu(y), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    return y + x


def _fusion(sd, p, x0, x1=None):
    """MiDaS FeatureFusionBlock_custom: x0 (+ resConfUnit1(x1)) -> resConfUnit2 -> x2 bilinear (align_corners) -> out_conv."""
    y = x0
    if x1 is not None:
        assert x0.shape == x1.shape
        y = y + _rcu(sd, p + "resConfUnit1.", x1)
    y = _rcu(sd, p + "resConfUnit2.", y)
    y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(y, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def dpt_neck(sd, feats, ph, pw, cfg=ZOED_N):
    """-> (fused [path_4 .. path_1], bottleneck = layer4_rn output)."""
    B = feats[0].shape[0]
    DIM = cfg["dim"]
    maps = []
    for i, t in enumerate(feats):
        p = f"{PP}act_postprocess{i + 1}."
        tok, cls = t[:, 1:], t[:, :1]
        y = torch.cat([tok, cls.expand_as(tok)], dim=-1)
        y = F.gelu(F.linear(y, sd[p + "0.project.0.weight"], sd[p + "0.project.0.bias"]))
        y = y.permute(0, 2, 1).reshape(B, DIM, ph, pw)
        y = F.conv2d(y, sd[p + "3.weight"], sd[p + "3.bias"])
        if i == 0:
            y = F.conv_transpose2d(y, sd[p + "4.weight"], sd[p + "4.bias"], stride=4)
        elif i == 1:
            y = F.conv_transpose2d(y, sd[p + "4.weight"], sd[p + "4.bias"], stride=2)
        elif i == 3:
            y = F.conv2d(y, sd[p + "4.weight"], sd[p + "4.bias"], stride=2, padding=1)
        maps.append(F.conv2d(y, sd[f"{SC}layer{i + 1}_rn.weight"], None, padding=1))
    p4 = _fusion(sd, SC + "refinenet4.", maps[3])
    p3 = _fusion(sd, SC + "refinenet3.", p4, maps[2])
    p2 = _fusion(sd, SC + "refinenet2.", p3, maps[1])
    p1 = _fusion(sd, SC + "refinenet1.", p2, maps[0])
    return [p4, p3, p2, p1], maps[3]


def relative_head(sd, p1):
    """scratch.output_conv: conv3x3 -> x2 bilinear (align_corners) -> conv3x3 + ReLU (= the 32-channel activation the bins
    head is conditioned on) -> conv1x1 + ReLU (= relative depth)."""
    p = SC + "output_conv."
    y = F.conv2d(p1, sd[p + "0.weight"], sd[p + "0.bias"], padding=1)
    y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
    act = F.relu(F.conv2d(y, sd[p + "2.weight"], sd[p + "2.bias"], padding=1))
    rel = F.relu(F.conv2d(act, sd[p + "4.weight"], sd[p + "4.bias"]))
    return rel.squeeze(1), act


def _mlp2(sd, p, x, act2=None):
    y = F.relu(F.conv2d(x, sd[p + "_net.0.weight"], sd[p + "_net.0.bias"]))
    y = F.conv2d(y, sd[p + "_net.2.weight"], sd[p + "_net.2.bias"])
    return act2(y) if act2 is not None else y


def inv_attractor(dx, alpha=300.0, gamma=2):
    return dx / (1 + alpha * dx.pow(gamma))


def log_binom(n, k, eps=1e-7):
    n = n + eps
    k = k + eps
    return n * torch.log(n) - k * torch.log(k) - (n - k) * torch.log(n - k + eps)


def metric_head(sd, act, bottleneck, blocks, rel, internals=None):
    x = F.conv2d(bottleneck, sd["conv2.weight"], sd["conv2.bias"])
    prev_bin = _mlp2(sd, "seed_bin_regressor.", x, F.softplus).float()        # SeedBinRegressorUnnormed
    prev_emb = _mlp2(sd, "seed_projector.", x)
    for i, feat in enumerate(blocks):
        emb = _mlp2(sd, f"projectors.{i}.", feat)
        # AttractorLayerUnnormed.forward
        y = emb + F.interpolate(prev_emb, emb.shape[-2:], mode="bilinear", align_corners=True)
        a = _mlp2(sd, f"attractors.{i}.", y, F.softplus).float()
        c = F.interpolate(prev_bin, a.shape[-2:], mode="bilinear", align_corners=True)
        delta = torch.zeros_like(c)
        for j in range(N_ATTRACTORS[i]):
            delta = delta + inv_attractor(a[:, j:j + 1] - c)
        delta = delta / N_ATTRACTORS[i]                                        # attractor_kind "mean"
        prev_bin = c + delta
        prev_emb = emb
        if internals is not None:
            internals.setdefault("bins", []).append(prev_bin)
    last = torch.cat([act, F.interpolate(rel.unsqueeze(1), size=act.shape[2:], mode="bilinear", align_corners=True).to(act.dtype)], dim=1)
    emb = F.interpolate(prev_emb, last.shape[-2:], mode="bilinear", align_corners=True)
    # ConditionalLogBinomial (bottleneck_factor 2, p_eps 1e-4, act softmax)
    y = F.gelu(F.conv2d(torch.cat([last, emb], dim=1), sd["conditional_log_binomial.mlp.0.weight"], sd["conditional_log_binomial.mlp.0.bias"]))
    pt = F.softplus(F.conv2d(y, sd["conditional_log_binomial.mlp.2.weight"], sd["conditional_log_binomial.mlp.2.bias"])).float()
    p = pt[:, :2] + 1e-4
    p = p[:, 0] / (p[:, 0] + p[:, 1])
    tmp = pt[:, 2:] + 1e-4
    tmp = (tmp[:, 0] / (tmp[:, 0] + tmp[:, 1])).unsqueeze(1)
    tmp = (MAX_TEMP - MIN_TEMP) * tmp + MIN_TEMP
    if internals is not None:
        internals["p"], internals["temperature"] = p, tmp
    p = p.unsqueeze(1)
    k = torch.arange(0, N_BINS, device=p.device, dtype=torch.float32).view(1, -1, 1, 1)
    km1 = torch.tensor([N_BINS - 1], device=p.device, dtype=torch.float32).view(1, -1, 1, 1)
    one_minus = torch.clamp(1 - p, 1e-4, 1)
    pc = torch.clamp(p, 1e-4, 1)
    yk = log_binom(km1, k) + k * torch.log(pc) + (km1 - k) * torch.log(one_minus)
    prob = torch.softmax(yk / tmp, dim=1)
    centers = F.interpolate(prev_bin, prob.shape[-2:], mode="bilinear", align_corners=True)
    return torch.sum(prob * centers, dim=1, keepdim=True)


def zoedepth_forward(sd, x, cfg=ZOED_N, return_all=False):
    """x: B,3,H,W normalised ((x - 0.5) / 0.5, zoedepth_model.py:79-83), H, W % 32 == 0 -> metric depth B,1,H,W
    (= model(x)['metric_depth'])."""
    B, _, H, W = x.shape
    assert H % 32 == 0 and W % 32 == 0
    ph, pw = H // PATCH, W // PATCH
    feats = beit_features(sd, x, cfg)
    blocks, bottleneck = dpt_neck(sd, feats, ph, pw, cfg)
    rel, act = relative_head(sd, blocks[-1])
    internals = {} if return_all else None
    out = metric_head(sd, act, bottleneck, blocks, rel, internals)
    if return_all:
        return {"feats": feats, "blocks": blocks, "bottleneck": bottleneck, "rel": rel, "act": act, "metric_depth": out, **internals}
    return out


def batch_infer(sd, im, flip_aug=True, edge_dilation=0, h_height=384, v_height=512, autocast_device=None, cfg=ZOED_N):
    """iw3/zoedepth_model.py:89-148 on top of zoedepth_forward (the preprocessing and dilation oracles are
    oracle/frames.py and oracle/iw3.py).  im: B,3,H,W in [0, 1] -> B,1,h,w (negated metric depth)."""
    from . import frames as ofr
    from . import iw3 as oiw3
    import numpy as np
    y, pad_h, pad_w = ofr.zoe_batch_preprocess(im.numpy() if torch.is_tensor(im) else im, h_height, v_height)
    x = torch.from_numpy(np.ascontiguousarray(y)).float()
    if flip_aug:
        x = torch.cat([x, torch.flip(x, dims=[3])], dim=0)
    if autocast_device is not None:
        x = x.to(autocast_device)
        sdd = {k: v.to(autocast_device) for k, v in sd.items()}
        with torch.autocast(device_type="cuda", dtype=torch.float16):
            out = zoedepth_forward(sdd, x, cfg)
        out = out.float().cpu()
    else:
        out = zoedepth_forward(sd, x, cfg)
    out = torch.nan_to_num(out)
    out = out[:, :, pad_h:-pad_h, pad_w:-pad_w]
    if edge_dilation:
        out = oiw3.dilate_edge(-out, edge_dilation)
    else:
        out = -out
    if flip_aug:
        n = out.shape[0] // 2
        out = (out[:n] + torch.flip(out[n:], dims=[3])) * 0.5
    return out
