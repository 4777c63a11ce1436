"""ORACLE (test infrastructure only - never imported by nunif_b200/): CPU restatement of the frame-edge
conversions (nunif/utils/video.py:218-223,236-246; iw3/utils.py:274-289) and of DepthAnything's
batch_preprocess (iw3/depth_anything_model.py:69-110), including ATen's antialiased bilinear resize
(aten/src/ATen/native/cpu/UpSampleKernel.cpp, _upsample_bilinear2d_aa; published algorithm restated).

Pinned against the real reference: tests/golden/frames.npz (oracle/gen_golden.py frames).
"""
import numpy as np


def hwc_to_chw_float(x):
    """uint8/uint16 (B)HWC -> float32 (B)CHW = x / iinfo.max (iw3/utils.py:285-287)."""
    maxv = np.float32(np.iinfo(x.dtype).max)
    perm = (2, 0, 1) if x.ndim == 3 else (0, 3, 1, 2)
    return np.ascontiguousarray(np.transpose(x, perm)).astype(np.float32) / maxv


def chw_float_to_hwc(x, use_16bit=False):
    """float (B)CHW -> uint8/uint16 (B)HWC = round_half_even(x * scale) (video.py:236-246)."""
    scale = np.float32(65535.0 if use_16bit else 255.0)
    perm = (1, 2, 0) if x.ndim == 3 else (0, 2, 3, 1)
    v = np.rint(np.transpose(x.astype(np.float32), perm) * scale)
    return np.clip(v, 0, scale).astype(np.uint16 if use_16bit else np.uint8)


def preprocess_size(H, W, lower_bound=392, max_aspect_ratio=4, limit_resolution=False):
    """depth_anything_model.py:69-101."""
    mult = 14
    if limit_resolution and lower_bound > min(W, H):
        lower_bound = min(W, H)
        lower_bound -= lower_bound % mult
        lower_bound = max(lower_bound, 224)
    scale_factor = lower_bound / W if W < H else lower_bound / H
    new_h, new_w = int(H * scale_factor), int(W * scale_factor)
    if new_h < new_w:
        new_w = min(new_w, int(max_aspect_ratio * new_h))
    else:
        new_h = min(new_h, int(max_aspect_ratio * new_w))
    new_h -= new_h % mult
    new_w -= new_w % mult
    return max(new_h, lower_bound), max(new_w, lower_bound)


def _aa_weights(in_size, out_size):
    """Per output index: (xmin, weights[]) of the antialiased triangle filter, align_corners=False."""
    f = np.float32
    scale = f(in_size) / f(out_size)
    support = scale if scale >= 1 else f(1)
    inv = f(1) / scale if scale >= 1 else f(1)
    taps = []
    for i in range(out_size):
        center = scale * (f(i) + f(0.5))
        xmin = max(0, int(center - support + f(0.5)))
        xsize = min(in_size, int(center + support + f(0.5))) - xmin
        w = np.array([max(f(0), f(1) - abs((f(j + xmin) - center + f(0.5)) * inv)) for j in range(xsize)], dtype=np.float32)
        taps.append((xmin, w / w.sum(dtype=np.float32)))
    return taps


def resize_bilinear_aa(x, new_h, new_w):
    """F.interpolate(x, (new_h, new_w), mode='bilinear', align_corners=False, antialias=True), x (..., H, W) fp32.
    Separable: horizontal pass, then vertical."""
    x = np.asarray(x, dtype=np.float32)
    H, W = x.shape[-2:]
    tw, th = _aa_weights(W, new_w), _aa_weights(H, new_h)
    hpass = np.zeros(x.shape[:-1] + (new_w,), dtype=np.float32)
    for ox, (xmin, w) in enumerate(tw):
        hpass[..., ox] = (x[..., xmin:xmin + len(w)] * w).sum(-1, dtype=np.float32)
    out = np.zeros(x.shape[:-2] + (new_h, new_w), dtype=np.float32)
    for oy, (ymin, w) in enumerate(th):
        out[..., oy, :] = (hpass[..., ymin:ymin + len(w), :] * w[:, None]).sum(-2, dtype=np.float32)
    return out


def batch_preprocess(x, lower_bound=392, max_aspect_ratio=4, limit_resolution=False):
    """depth_anything_model.py:69-110 for BCHW float32 in [0,1]."""
    B, C, H, W = x.shape
    nh, nw = preprocess_size(H, W, lower_bound, max_aspect_ratio, limit_resolution)
    y = np.clip(resize_bilinear_aa(x, nh, nw), 0, 1)
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32).reshape(1, 3, 1, 1)
    stdv = np.array([0.229, 0.224, 0.225], dtype=np.float32).reshape(1, 3, 1, 1)
    return (y - mean) / stdv


def zoe_preprocess_size(H, W, h_height=384, v_height=512, mod=32):
    """iw3/zoedepth_model.py:30-71 -> (new_h, new_w, pad_h, pad_w, frame_h, frame_w)."""
    target = h_height if W > H else v_height
    if target < H:
        new_h = target
        new_w = int(new_h / H * W)
        if new_w % mod != 0:
            new_w += mod - new_w % mod
        if new_h % mod != 0:
            new_h += mod - new_h % mod
    else:
        new_h, new_w = H, W
        new_w -= new_w % mod
        new_h -= new_h % mod
    pad_src_h, pad_src_w = int((H * 0.5) ** 0.5 * 3), int((W * 0.5) ** 0.5 * 3)
    sh, sw = pad_src_h / (H + pad_src_h * 2), pad_src_w / (W + pad_src_w * 2)
    if new_h > new_w:
        pad_h = round(new_h * sh)
        frame_h = new_h - pad_h * 2
        frame_w = int(W * (frame_h / H))
        frame_w += frame_w % 2
        pad_w = (new_h - frame_w) // 2
    else:
        pad_h, pad_w = round(new_h * sh), round(new_w * sw)
        frame_h, frame_w = new_h - pad_h * 2, new_w - pad_w * 2
    return new_h, new_w, pad_h, pad_w, frame_h, frame_w


def zoe_batch_preprocess(x, h_height=384, v_height=512, mod=32):
    """zoedepth_model.py:30-85: AA resize, reflection pad (reflection_pad2d_loop == periodic mirror extension, which is
    what numpy's mode="reflect" produces for pads larger than the image), clamp, (x - 0.5) / 0.5."""
    H, W = x.shape[-2:]
    _, _, ph, pw, fh, fw = zoe_preprocess_size(H, W, h_height, v_height, mod)
    y = resize_bilinear_aa(x, fh, fw)
    y = np.pad(y, ((0, 0), (0, 0), (ph, ph), (pw, pw)), mode="reflect")
    y = np.clip(y, 0, 1)
    return ((y - np.float32(0.5)) / np.float32(0.5)).astype(np.float32), ph, pw
