"""ORACLE (test infrastructure only - never imported by nunif_b200/): CPU restatement of iw3's output composition:
the red-cyan anaglyph family (iw3/anaglyph.py:4-110), TF.resize(BICUBIC, antialias=True) = ATen
_upsample_bicubic2d_aa, and postprocess_image / postprocess_padding (iw3/utils.py:394-487, VR180 excluded).

Pinned against the real reference: tests/golden/anaglyph.npz (all methods) and tests/golden/postprocess.npz, produced by
executing the reference's own postprocess_image (oracle/gen_golden.py postprocess).
"""
import numpy as np
from .swin_unet import bicubic_aa_weights


def _gray601(x):
    return x[0:1] * np.float32(0.299) + x[1:2] * np.float32(0.587) + x[2:3] * np.float32(0.114)


def anaglyph(l, r, kind):
    """l, r: (3,H,W) float32 numpy."""
    f = np.float32
    if kind == "color":
        return np.concatenate([l[0:1], r[1:3]], 0)
    if kind == "half-color":
        return np.clip(np.concatenate([_gray601(l), r[1:3]], 0), 0, 1)
    if kind == "gray":
        ry = _gray601(r)
        return np.clip(np.concatenate([_gray601(l), ry, ry], 0), 0, 1)
    if kind == "wimmer":
        return np.clip(np.concatenate([l[1:2] * f(0.7) + l[2:3] * f(0.3), r[1:3]], 0), 0, 1)
    if kind == "wimmer2":
        g_l = l[1:2] + f(0.45) * np.maximum(l[0:1] - l[1:2], 0)
        b_l = l[2:3] + f(0.25) * np.maximum(l[0:1] - l[2:3], 0)
        g_r = r[1:2] + f(0.45) * np.maximum(r[0:1] - r[1:2], 0)
        b_r = r[2:3] + f(0.25) * np.maximum(r[0:1] - r[2:3], 0)
        left = np.power(f(0.75) * g_l + f(0.25) * b_l, f(1.0 / 1.6))
        return np.clip(np.concatenate([left, g_r, b_r], 0), 0, 1).astype(np.float32)
    raise ValueError(kind)


def resize_bicubic_aa(x, oh, ow):
    """(..., H, W) float32 -> (..., oh, ow): horizontal pass then vertical pass with the ATen AA weights."""
    x = np.asarray(x, dtype=np.float32)
    H, W = x.shape[-2:]
    sx, wx = bicubic_aa_weights(W, ow)
    sy, wy = bicubic_aa_weights(H, oh)
    sx, wx, sy, wy = sx.numpy(), wx.numpy(), sy.numpy(), wy.numpy()
    hp = np.zeros(x.shape[:-1] + (ow,), dtype=np.float32)
    for o in range(ow):
        n = min(wx.shape[1], W - sx[o])
        hp[..., o] = (x[..., sx[o]:sx[o] + n] * wx[o, :n]).sum(-1, dtype=np.float32)
    out = np.zeros(x.shape[:-2] + (oh, ow), dtype=np.float32)
    for o in range(oh):
        n = min(wy.shape[1], H - sy[o])
        out[..., o, :] = (hp[..., sy[o]:sy[o] + n, :] * wy[o, :n, None]).sum(-2, dtype=np.float32)
    return out


def _pad(x, left, top, right, bottom):
    return np.pad(x, ((0, 0), (top, bottom), (left, right)), mode="constant")


def postprocess_image(left, right, ipd_offset=0, pad=None, pad_mode=None, half_sbs=False, half_tb=False, tb=False, cross_eyed=False,
                      anaglyph_type=None, max_output_height=None, max_output_width=None, keep_aspect_ratio=False, dubois=None):
    """iw3/utils.py:430-487 on (3,H,W) float32 numpy arrays; `dubois` = callable for the dubois methods (oracle.iw3.dubois)."""
    ipd_pad = int(abs(ipd_offset) * 0.01 * max(left.shape[-2:]))
    ipd_pad -= ipd_pad % 2
    if ipd_pad > 0:
        po, pi = (ipd_pad * 2, ipd_pad) if ipd_offset > 0 else (ipd_pad, ipd_pad * 2)
        left, right = _pad(left, po, 0, pi, 0), _pad(right, pi, 0, po, 0)
    if pad is not None or pad_mode == "16:9":
        if pad_mode in {"tblr", "tb", "lr"}:
            ph = round(left.shape[1] * pad) // 2 if "tb" in pad_mode else 0
            pw = round(left.shape[2] * pad) // 2 if "lr" in pad_mode else 0
            left, right = _pad(left, pw, ph, pw, ph), _pad(right, pw, ph, pw, ph)
        elif pad_mode == "top":
            pt = round(left.shape[1] * pad)
            left, right = _pad(left, 0, pt, 0, 0), _pad(right, 0, pt, 0, 0)
        else:
            h, w = left.shape[1:]
            if abs(16 / 9 - w / h) > 1e-3:
                ph = pw = 0
                if w / h > 16 / 9:
                    ph = (round(w / (16 / 9)) - h) // 2
                else:
                    pw = (round(h * (16 / 9)) - w) // 2
                left, right = _pad(left, pw, ph, pw, ph), _pad(right, pw, ph, pw, ph)
    if half_sbs:
        left, right = resize_bicubic_aa(left, left.shape[1], left.shape[2] // 2), resize_bicubic_aa(right, right.shape[1], right.shape[2] // 2)
    elif half_tb:
        left, right = resize_bicubic_aa(left, left.shape[1] // 2, left.shape[2]), resize_bicubic_aa(right, right.shape[1] // 2, right.shape[2])
    if anaglyph_type is not None:
        sbs = dubois(left, right, anaglyph_type == "dubois") if anaglyph_type in {"dubois", "dubois2"} else anaglyph(left, right, anaglyph_type)
    elif tb or half_tb:
        sbs = np.clip(np.concatenate([left, right], 1), 0, 1)
    elif cross_eyed:
        sbs = np.clip(np.concatenate([right, left], 2), 0, 1)
    else:
        sbs = np.clip(np.concatenate([left, right], 2), 0, 1)
    h, w = sbs.shape[1:]
    nw, nh = w, h
    if max_output_height is not None and nh > max_output_height:
        if keep_aspect_ratio:
            nw = int(max_output_height / nh * nw)
        nh = max_output_height
    if max_output_width is not None and nw > max_output_width:
        if keep_aspect_ratio:
            nh = int(max_output_width / nw * nh)
        nw = max_output_width
    if nw != w or nh != h:
        nh -= nh % 2
        nw -= nw % 2
        sbs = np.clip(resize_bicubic_aa(sbs, nh, nw), 0, 1)
    return sbs


def _grid_sample_bicubic_zeros(c, gx, gy):
    """F.grid_sample(mode="bicubic", padding_mode="zeros", align_corners=True) for one (C,H,W) image and (Ho,Wo) grids in
    [-1,1] (ATen grid_sampler_2d: A = -0.75, out-of-range taps contribute 0).  Explicit 4x4 gather."""
    f = np.float32
    C, H, W = c.shape
    ix = ((gx + f(1)) / f(2)) * f(W - 1)
    iy = ((gy + f(1)) / f(2)) * f(H - 1)
    ix0, iy0 = np.floor(ix), np.floor(iy)
    tx, ty = (ix - ix0).astype(np.float32), (iy - iy0).astype(np.float32)

    def coeffs(t):
        A = f(-0.75)
        def c1(x): return ((A + f(2)) * x - (A + f(3))) * x * x + f(1)
        def c2(x): return ((A * x - f(5) * A) * x + f(8) * A) * x - f(4) * A
        return [c2(t + f(1)), c1(t), c1(f(1) - t), c2(f(2) - t)]
    wx, wy = coeffs(tx), coeffs(ty)
    out = np.zeros((C,) + gx.shape, dtype=np.float32)
    for a in range(4):
        yy = (iy0 + a - 1).astype(np.int64)
        row = np.zeros_like(out)
        for b in range(4):
            xx = (ix0 + b - 1).astype(np.int64)
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            v = c[:, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * ok
            row = row + v * wx[b]
        out = out + row * wy[a]
    return out


def equirectangular_projection(c):
    """iw3/equirectangular.py:7-40 (VR180 output): zero-pad to a square of 1.5x the longer edge, then bicubic grid_sample
    through the tan() mesh.  Pinned for the round that ports it (the engine raises NotImplementedError for vr180 today)."""
    import math
    f = np.float32
    h, w = c.shape[1:]
    max_edge = max(h, w)
    output_size = max_edge + max_edge // 2
    pad_w, pad_h = (output_size - w) // 2, (output_size - h) // 2
    c = np.pad(np.asarray(c, dtype=np.float32), ((0, 0), (pad_h, pad_h), (pad_w, pad_w)))
    h, w = c.shape[1:]
    import torch   # linspace with torch's fp32 semantics (symmetric evaluation)
    y = torch.linspace(-1, 1, h).numpy()[:, None].repeat(w, 1)
    x = torch.linspace(-1, 1, w).numpy()[None, :].repeat(h, 0)
    az, el = x * f(math.pi * 0.5), y * f(math.pi * 0.5)
    k = f(max_edge / output_size)
    mesh_x = k * np.tan(az).astype(np.float32)
    mesh_y = k * (np.tan(el).astype(np.float32) / np.cos(az).astype(np.float32))
    return np.clip(_grid_sample_bicubic_zeros(c, mesh_x.astype(np.float32), mesh_y.astype(np.float32)), 0, 1)
