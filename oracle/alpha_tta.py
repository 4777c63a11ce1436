"""ORACLE (test infrastructure only - never imported by nunif_b200/): CPU restatement of
AlphaBorderPadding (nunif/utils/alpha.py:32-57) and tta_split / tta_merge (nunif/transforms/tta.py:20-48).

Pinned against outputs of the real reference: tests/golden/alpha_tta.npz (oracle/gen_golden.py alpha_tta).
numpy, explicit index arithmetic (no conv / flip / rot90 calls), so that it is an independent statement of
what the CUDA kernels in csrc/alpha_tta.cu must compute.
"""
import numpy as np


def _box3_zero(a):
    """3x3 box sum with zero padding over the last two axes (ChannelWiseSum, alpha.py:5-29), fp32 accumulation."""
    p = np.zeros(a.shape[:-2] + (a.shape[-2] + 2, a.shape[-1] + 2), dtype=np.float32)
    p[..., 1:-1, 1:-1] = a
    H, W = a.shape[-2:]
    s = np.zeros_like(a, dtype=np.float32)
    for dy in range(3):
        for dx in range(3):
            s = s + p[..., dy:dy + H, dx:dx + W]
    return s


def alpha_border_padding(rgb, alpha, offset):
    """rgb (3,H,W), alpha (1,H,W) float32 -> (3,H,W).  alpha.py:39-56."""
    rgb = np.array(rgb, dtype=np.float32, copy=True)
    a = np.asarray(alpha, dtype=np.float32)[0]
    mask = (a > 0).astype(np.float32)                    # :43-44
    nega = mask < 1
    rgb[:, nega] = 0                                     # :45-46
    for _ in range(offset):
        mw = _box3_zero(mask)                            # :48
        border = _box3_zero(rgb) / (mw + np.float32(1e-7))   # :49-50
        rgb[:, nega] = border[:, nega]                   # :51
        mask = (mw > 0).astype(np.float32)               # :52-53
        nega = mask < 1
    return np.clip(rgb, 0, 1)                            # :56


def tta_view(x, k):
    """View k of tta_split (tta.py:20-33) by index arithmetic: out[c,i,j] = x[c, sy, sx]."""
    C, H, W = x.shape
    oh, ow = (H, W) if k < 4 else (W, H)
    i, j = np.meshgrid(np.arange(oh), np.arange(ow), indexing="ij")
    sy, sx = [
        (i, j), (i, W - 1 - j), (H - 1 - i, j), (H - 1 - i, W - 1 - j),
        (j, W - 1 - i), (H - 1 - j, W - 1 - i), (j, i), (H - 1 - j, i),
    ][k]
    return x[:, sy, sx]


def tta_split(x):
    return tuple(tta_view(x, k) for k in range(8))


def tta_merge(zs):
    """tta.py:36-48: sum of the inverse-transformed views in order, * 1/8, clamp."""
    C, H, W = zs[0].shape
    y, x = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    pos = [
        (y, x), (y, W - 1 - x), (H - 1 - y, x), (H - 1 - y, W - 1 - x),
        (W - 1 - x, y), (W - 1 - x, H - 1 - y), (x, y), (x, H - 1 - y),
    ]
    acc = np.array(zs[0][:, pos[0][0], pos[0][1]], dtype=np.float32)
    for k in range(1, 8):
        acc = acc + np.asarray(zs[k], dtype=np.float32)[:, pos[k][0], pos[k][1]]
    acc = acc * np.float32(1 / 8.0)
    return np.clip(acc, 0, 1)
