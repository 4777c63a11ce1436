"""Oracle: iw3 depth post-processing and stereo warps (TEST INFRASTRUCTURE).

Plain torch-CPU / numpy fp32 restatements of
  iw3/backward_warp.py:67-121   (apply_divergence_grid_sample)
  iw3/forward_warp.py:18-256    (depth_order_bilinear_forward_warp)
  iw3/dilation.py:30-142        (dilate_edge)
  iw3/depth_scaler.py:4-17      (minmax_normalize)
  iw3/mapper.py:29-32           (div_* mappers)
  iw3/anaglyph.py:51-110        (dubois & friends)
  iw3/utils.py:460-469          (SBS compose)
The torch library ops the reference leans on (grid_sample, max_pool2d,
index_copy_) are restated with explicit index arithmetic so the oracle is an
independent statement; oracle/gen_golden.py pins it to the real reference.
"""
import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------

def bilinear_resize_align_corners(x, out_h, out_w):
    """F.interpolate(mode='bilinear', align_corners=True, antialias=False)
    (used at backward_warp.py:70-71) with explicit gathers."""
    B, C, h, w = x.shape
    if (h, w) == (out_h, out_w):
        return x
    def axis(n_in, n_out):
        scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        src = torch.arange(n_out, dtype=torch.float32) * torch.tensor(scale, dtype=torch.float32)
        i0 = src.floor().long().clamp(0, n_in - 1)
        i1 = (i0 + 1).clamp(max=n_in - 1)
        t = src - i0.float()
        return i0, i1, t
    y0, y1, ty = axis(h, out_h)
    x0, x1, tx = axis(w, out_w)
    top = x[:, :, y0][:, :, :, x0] * (1 - tx) + x[:, :, y0][:, :, :, x1] * tx
    bot = x[:, :, y1][:, :, :, x0] * (1 - tx) + x[:, :, y1][:, :, :, x1] * tx
    ty = ty.view(1, 1, -1, 1)
    return top * (1 - ty) + bot * ty


def grid_sample_bilinear_border(c, grid):
    """F.grid_sample(mode='bilinear', padding_mode='border', align_corners=True)
    (backward_warp.py:81).  grid: B,H,W,2 in [-1,1]."""
    B, C, H, W = c.shape
    gx = (grid[..., 0] + 1) * 0.5 * (W - 1)
    gy = (grid[..., 1] + 1) * 0.5 * (H - 1)
    gx = gx.clamp(0, W - 1)
    gy = gy.clamp(0, H - 1)
    x0 = gx.floor()
    y0 = gy.floor()
    tx = gx - x0
    ty = gy - y0
    x0 = x0.long()
    y0 = y0.long()
    x1 = (x0 + 1).clamp(max=W - 1)
    y1 = (y0 + 1).clamp(max=H - 1)
    out = torch.empty((B, C, grid.shape[1], grid.shape[2]), dtype=c.dtype)
    for b in range(B):
        cb = c[b]
        v00 = cb[:, y0[b], x0[b]]
        v01 = cb[:, y0[b], x1[b]]
        v10 = cb[:, y1[b], x0[b]]
        v11 = cb[:, y1[b], x1[b]]
        out[b] = (v00 * ((1 - tx[b]) * (1 - ty[b])) + v01 * (tx[b] * (1 - ty[b]))
                  + v10 * ((1 - tx[b]) * ty[b]) + v11 * (tx[b] * ty[b]))
    return out


# ---------------------------------------------------------------------------
# backward warp
# ---------------------------------------------------------------------------

def make_grid(batch, width, height):
    """backward_warp.py:86-93."""
    my, mx = torch.meshgrid(torch.linspace(-1, 1, height), torch.linspace(-1, 1, width), indexing="ij")
    my = my.reshape(1, 1, height, width).expand(batch, 1, height, width)
    mx = mx.reshape(1, 1, height, width).expand(batch, 1, height, width)
    return torch.cat((mx, my), dim=1)


def backward_warp(c, grid, delta, delta_scale):
    """backward_warp.py:67-83."""
    grid = grid + delta * delta_scale
    grid = bilinear_resize_align_corners(grid, c.shape[2], c.shape[3])
    z = grid_sample_bilinear_border(c, grid.permute(0, 2, 3, 1))
    return torch.clamp(z, 0, 1)


def apply_divergence_grid_sample(c, depth, divergence, convergence, synthetic_view="both"):
    """backward_warp.py:96-121."""
    B, _, H, W = depth.shape
    if synthetic_view != "both":
        divergence = divergence * 2
    base_size = max(H, W)
    shift_size = divergence * 0.01
    index_shift = depth * shift_size - (shift_size * convergence)
    delta = torch.cat([index_shift, torch.zeros_like(index_shift)], dim=1)
    delta_scale = base_size / W
    grid = make_grid(B, W, H)
    if synthetic_view == "both":
        return backward_warp(c, grid, -delta, delta_scale), backward_warp(c, grid, delta, delta_scale)
    if synthetic_view == "right":
        return c, backward_warp(c, grid, delta, delta_scale)
    return backward_warp(c, grid, -delta, delta_scale), c


# ---------------------------------------------------------------------------
# forward warp
# ---------------------------------------------------------------------------

def _shift_fill(x, sign, max_tries=100):
    """forward_warp.py:18-30 (flip_sign=False path)."""
    mask = x < 0
    while bool(mask.any()) and max_tries > 0:
        if sign > 0:
            nb = F.pad(x[:, :, :, 1:], (0, 1, 0, 0))
        else:
            nb = F.pad(x[:, :, :, :-1], (1, 0, 0, 0))
        x = torch.where(mask, nb, x)
        mask = x < 0
        max_tries -= 1
    return x


def _shift_fill_pack(left, right):
    """forward_warp.py:33-42 (inconsistent_shift=False)."""
    left = _shift_fill(left, -1)
    right = torch.flip(_shift_fill(torch.flip(right, dims=(-1,)), -1), dims=(-1,))
    return left, right


def _fix_layered_holes(side, index, sign, max_tries=100):
    """forward_warp.py:45-59.  Returns new (side, index)."""
    def mk(idx):
        d = (idx[:, :, :, :-1] - idx[:, :, :, 1:]) > 0
        return F.pad(d, (0, 1, 0, 0)) if sign > 0 else F.pad(d, (1, 0, 0, 0))
    side = side.clone()
    index = index.clone()
    mask = mk(index)
    while bool(mask.any()) and max_tries > 0:
        side = torch.where(mask.expand_as(side), torch.full_like(side, -2.0), side)
        if sign > 0:
            nb = F.pad(index[:, :, :, 1:], (0, 1, 0, 0))
        else:
            nb = F.pad(index[:, :, :, :-1], (1, 0, 0, 0))
        index = torch.where(mask, nb, index)
        mask = mk(index)
        max_tries -= 1
    return side, index


def _warp(c5, index_shift, depth):
    """forward_warp.py:75-132: bilinear splat in ascending-depth order, last
    writer wins (numpy fancy assignment is sequential => same rule)."""
    B, CH, H, W = c5.shape
    xs = torch.arange(0, W).view(1, 1, W).expand(B, H, W)
    float_index = torch.clamp(xs + index_shift, 0, W - 1)
    floor_index = torch.clamp(float_index.floor(), 0, W - 1)
    ceil_index = torch.clamp(float_index.ceil(), 0, W - 1)
    ceil_w = torch.clamp(float_index - floor_index, min=1e-5, max=1.0 - 1e-5)
    floor_w = 1.0 - ceil_w
    row_base = (torch.arange(H).view(1, H, 1) * W + torch.arange(B).view(B, 1, 1) * H * W)
    fl = (floor_index.long() + row_base).reshape(-1).numpy()
    ce = (ceil_index.long() + row_base).reshape(-1).numpy()
    order = np.argsort(depth.reshape(-1).numpy(), kind="stable")
    data = c5.permute(0, 2, 3, 1).reshape(-1, CH).numpy()
    fdat = np.concatenate([floor_w.reshape(-1, 1).numpy(), data], axis=1)
    cdat = np.concatenate([ceil_w.reshape(-1, 1).numpy(), data], axis=1)
    undef = np.array([0.0] + [-1.0] * CH, dtype=np.float32)
    fout = np.tile(undef, (data.shape[0], 1))
    cout = np.tile(undef, (data.shape[0], 1))
    fout[fl[order]] = fdat[order]
    cout[ce[order]] = cdat[order]
    fw, fv = fout[:, 0:1], fout[:, 1:]
    cw, cv = cout[:, 0:1], cout[:, 1:]
    with np.errstate(invalid="ignore", divide="ignore"):
        out = (fv * fw + cv * cw) / (fw + cw)
    out = np.nan_to_num(out, nan=-1.0)
    return torch.from_numpy(out.astype(np.float32)).view(B, H, W, CH).permute(0, 3, 1, 2)


def upsample_depth(depth, size):
    """forward_warp.py:146-148.  The antialiased bilinear resize is an ATen
    library op; the oracle calls it (restating ATen's separable AA kernel
    bit-exactly is not possible from the reference tree)."""
    if tuple(depth.shape[-2:]) == tuple(size):
        return depth
    return F.interpolate(depth, size=size, mode="bilinear", align_corners=True, antialias=True)


def forward_warp(c, depth, divergence, convergence, fill=True, synthetic_view="both",
                 return_mask=False, width_base=True):
    """forward_warp.py:140-243 (inconsistent_shift=False)."""
    src = c
    depth = upsample_depth(depth, c.shape[-2:])
    if synthetic_view != "both":
        divergence = divergence * 2
    base = c.shape[-1] if width_base else max(c.shape[-2:])
    P = int(base * divergence * 0.01 + 2)
    c = F.pad(c, (P, P, 0, 0), mode="replicate")
    depth = F.pad(depth, (P, P, 0, 0), mode="replicate")
    B, _, H, W = depth.shape
    shift_size = divergence * 0.01 * base * 0.5
    index_shift = (depth * shift_size - (shift_size * convergence)).view(B, H, W)
    xidx = torch.arange(0, W).view(1, 1, 1, W).expand(B, 1, H, W).to(c.dtype)
    c5 = torch.cat([c, xidx], dim=1)

    def gen_mask2(m):
        m = m[:, 0:1]
        return torch.clamp((m == -1).float() + (m == -2).float() * 0.5, 0, 1)

    left = right = left_mask = right_mask = None
    if synthetic_view in ("both", "left"):
        e = _warp(c5, index_shift, depth)[:, :, :, P:W - P]
        left, left_idx = e[:, :-1], e[:, -1:]
    if synthetic_view in ("both", "right"):
        e = _warp(c5, -index_shift, depth)[:, :, :, P:W - P]
        right, right_idx = e[:, :-1], e[:, -1:]
    if left is not None:
        left_idx = _shift_fill(left_idx, -1)
        left, _ = _fix_layered_holes(left, left_idx, 1)
        left_mask = gen_mask2(left)
        left = _shift_fill(left, -1) if fill else torch.clamp(left, 0, 1)
    if right is not None:
        right_idx = _shift_fill(right_idx, 1)
        right, _ = _fix_layered_holes(right, right_idx, -1)
        right_mask = gen_mask2(right)
        right = _shift_fill(right, 1) if fill else torch.clamp(right, 0, 1)
    if left is None:
        left = src
    if right is None:
        right = src
    if return_mask:
        return left.contiguous(), right.contiguous(), left_mask, right_mask
    return left.contiguous(), right.contiguous()


# ---------------------------------------------------------------------------
# dilation / normalisation / mapper
# ---------------------------------------------------------------------------

def _maxpool(x, kh, kw):
    """F.max_pool2d(kernel=(kh,kw), stride=1, padding=k//2) (dilation.py:41-46):
    -inf padding, explicit shifted maxima."""
    ph, pw = kh // 2, kw // 2
    xp = F.pad(x, (pw, pw, ph, ph), value=float("-inf"))
    H, W = x.shape[-2:]
    out = None
    for dy in range(kh):
        for dx in range(kw):
            v = xp[:, :, dy:dy + H, dx:dx + W]
            out = v if out is None else torch.maximum(out, v)
    return out


def gaussian_blur(x):
    """dilation.py:30-38."""
    k = torch.tensor([[21, 31, 21], [31, 48, 31], [21, 31, 21]], dtype=torch.float32).reshape(1, 1, 3, 3) / 256.0
    return F.conv2d(F.pad(x, [1] * 4, mode="replicate"), k)


def edge_weight(x):
    """dilation.py:101-112."""
    max_v = _maxpool(x, 3, 3)
    min_v = -_maxpool(-x, 3, 3)
    range_v = max_v - min_v
    range_c = range_v - range_v.mean(dim=[1, 2, 3], keepdim=True)
    range_s = range_c.pow(2).mean(dim=[1, 2, 3], keepdim=True).sqrt()
    w = (range_c / (range_s + 1e-6)).clamp(-3, 3)
    w_min, w_max = w.amin(dim=[1, 2, 3], keepdim=True), w.amax(dim=[1, 2, 3], keepdim=True)
    return (w - w_min) / ((w_max - w_min) + 1e-6)


def edge_dilation_parse(n):
    """dilation.py:5-22."""
    if isinstance(n, (list, tuple)):
        if len(n) == 0:
            return 0, 0
        if len(n) == 1:
            return n[0], n[0]
        return n[0], n[1]
    if isinstance(n, int):
        return n, n
    if n is None:
        return 0, 0
    raise ValueError(f"Unsupported edge_dilation type {type(n)}. Supported types: int, list, tuple.")


def dilate_edge(x, n):
    """dilation.py:115-142."""
    x_iter, y_iter = edge_dilation_parse(n)
    xy = min(x_iter, y_iter)
    plan = [(3, 3)] * xy + [(3, 1)] * (y_iter - xy) + [(1, 3)] * (x_iter - xy)
    for kh, kw in plan:
        w = edge_weight(x)
        x2 = _maxpool(gaussian_blur(x), kh, kw)
        x = (x * (1 - w)) + (x2 * w)
    return x


def minmax_normalize(depth):
    """depth_scaler.py:4-17 with per-frame amin/amax (base_depth_model.py:176-194,
    default decay=0/buffer=1): depth B,1,h,w -> per-frame [0,1]."""
    out = []
    for d in depth:
        mn, mx = d.amin(), d.amax()
        scale = mx - mn
        out.append(((d - mn) / scale).clamp(0, 1) if scale > 0 else d.clamp(0, 1))
    return torch.stack(out)


def distance_to_disparity(x, c):
    """mapper.py:29-32."""
    c1 = 1.0 + c
    min_v = c / c1
    return ((c / (c1 - x)) - min_v) / (1.0 - min_v)


_DIV_C = {"div_25": 2.5, "div_10": 1.0, "div_6": 0.6, "div_4": 0.4, "div_2": 0.2, "div_1": 0.1}  # mapper.py:106-113


def mapper(x, name):
    """mapper.py:63-118 subset on the hot path: none / div_* (metric)."""
    if name == "none":
        return x
    if name in _DIV_C:
        return distance_to_disparity(x, _DIV_C[name])
    raise NotImplementedError(f"mapper={name}")


# ---------------------------------------------------------------------------
# compose
# ---------------------------------------------------------------------------

def dubois(left, right, clip_before=True):
    """anaglyph.py:51-92.  left/right: 3,H,W."""
    def to_linear(x):
        return torch.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)

    def to_nonlinear(x):
        return torch.where(x <= 0.0031308, x * 12.92, 1.055 * x ** (1.0 / 2.4) - 0.055)

    def dot_clip(x, vec):
        v = (x * vec).sum(dim=0, keepdim=True)
        return v.clamp(0, 1) if clip_before else v

    l, r = to_linear(left.clone()), to_linear(right.clone())
    l_mat = torch.tensor([[0.437, 0.449, 0.164], [-0.062, -0.062, -0.024], [-0.048, -0.050, -0.017]]).reshape(3, 3, 1, 1)
    r_mat = torch.tensor([[-0.011, -0.032, -0.007], [0.377, 0.761, 0.009], [-0.026, -0.093, 1.234]]).reshape(3, 3, 1, 1)
    a = torch.cat([dot_clip(l, l_mat[i]) + dot_clip(r, r_mat[i]) for i in range(3)], dim=0)
    a = torch.clamp(a, 0, 1)
    a = to_nonlinear(a)
    return torch.clamp(a, 0, 1)


def sbs(left, right):
    """iw3/utils.py:466-469: cat along width + clamp."""
    return torch.clamp(torch.cat([left, right], dim=-1), 0., 1.)
