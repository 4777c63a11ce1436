"""Oracle: SwinUNet (1x/2x/4x/downscaled) forward from a state_dict
(TEST INFRASTRUCTURE).

Functional fp32 restatement of waifu2x/models/swin_unet.py:119-199, :208-387
and of torchvision's ``shifted_window_attention`` / ``SwinTransformerBlock``
(torchvision/models/swin_transformer.py:116-229, :401-455; third-party, pinned
by execution against the container's torchvision 0.26 in oracle/gen_golden.py).
Norm layers are Identity (swin_unet.py:16-17); MLP = Linear-GELU-Linear with
ratio 2 (swin_unet.py:31).
"""
import torch
import torch.nn.functional as F

WS = 6


def window_attention(x, p, sd, heads, shift):
    """x: B,H,W,C.  torchvision swin_transformer.py:116-229 with window 6x6."""
    B, H, W, C = x.shape
    assert H % WS == 0 and W % WS == 0
    d = C // heads
    s = shift if (WS < H) else 0  # :151-155 (no shift when the window covers the map)
    if s > 0:
        x = torch.roll(x, shifts=(-s, -s), dims=(1, 2))
    nh, nw = H // WS, W // WS
    xw = x.view(B, nh, WS, nw, WS, C).permute(0, 1, 3, 2, 4, 5).reshape(B * nh * nw, WS * WS, C)
    qkv = F.linear(xw, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    qkv = qkv.reshape(xw.size(0), WS * WS, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * d ** -0.5
    attn = q.matmul(k.transpose(-2, -1))
    table = sd[p + ".relative_position_bias_table"]
    index = sd[p + ".relative_position_index"]
    bias = table[index].view(WS * WS, WS * WS, -1).permute(2, 0, 1).unsqueeze(0)
    attn = attn + bias
    if s > 0:
        # :193-209 region-id mask, -100 across regions
        m = x.new_zeros((H, W))
        slices = ((0, -WS), (-WS, -s), (-s, None))
        cnt = 0
        for hs in slices:
            for ws_ in slices:
                m[hs[0]:hs[1], ws_[0]:ws_[1]] = cnt
                cnt += 1
        m = m.view(nh, WS, nw, WS).permute(0, 2, 1, 3).reshape(nh * nw, WS * WS)
        m = m.unsqueeze(1) - m.unsqueeze(2)
        m = m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)
        attn = attn.view(B, nh * nw, heads, WS * WS, WS * WS) + m.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, WS * WS, WS * WS)
    attn = F.softmax(attn, dim=-1)
    o = attn.matmul(v).transpose(1, 2).reshape(xw.size(0), WS * WS, C)
    o = F.linear(o, sd[p + ".proj.weight"], sd[p + ".proj.bias"])
    o = o.view(B, nh, nw, WS, WS, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if s > 0:
        o = torch.roll(o, shifts=(s, s), dims=(1, 2))
    return o


def swin_blocks(x, p, sd, heads, layers):
    """swin_unet.py:20-42 + torchvision SwinTransformerBlock.forward :452-455."""
    for i in range(layers):
        bp = f"{p}.block.{i}"
        x = x + window_attention(x, bp + ".attn", sd, heads, 0 if i % 2 == 0 else WS // 2)
        h = F.gelu(F.linear(x, sd[bp + ".mlp.0.weight"], sd[bp + ".mlp.0.bias"]))
        x = x + F.linear(h, sd[bp + ".mlp.3.weight"], sd[bp + ".mlp.3.bias"])
    return x


def patch_down(x, p, sd):
    """swin_unet.py:45-62."""
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2)
    return x.permute(0, 2, 3, 1).contiguous()


def patch_up(x, p, sd):
    """swin_unet.py:65-82."""
    x = F.linear(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"])
    x = F.pixel_shuffle(x.permute(0, 3, 1, 2), 2)
    return x.permute(0, 2, 3, 1).contiguous()


def to_image(x, p, sd, scale_factor):
    """swin_unet.py:85-116."""
    x = F.linear(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"]).permute(0, 3, 1, 2)
    if scale_factor > 1:
        x = F.pixel_shuffle(x, scale_factor)
    return x


def swin_unet_base(sd, x, scale_factor, p="unet"):
    """SwinUNetBase.forward, swin_unet.py:180-199."""
    C = sd[p + ".patch.2.weight"].shape[0]
    heads = C // 16
    x2 = F.leaky_relu(F.conv2d(x, sd[p + ".patch.0.weight"], sd[p + ".patch.0.bias"]), 0.1)
    x2 = F.leaky_relu(F.conv2d(x2, sd[p + ".patch.2.weight"], sd[p + ".patch.2.bias"]), 0.1)
    x2 = x2[:, :, 6:-6, 6:-6]
    assert x2.shape[2] % 12 == 0 and x2.shape[2] % 16 == 0
    x2 = x2.permute(0, 2, 3, 1).contiguous()
    x3 = swin_blocks(x2, p + ".swin1", sd, heads, 2)
    x4 = patch_down(x3, p + ".down1", sd)
    x4 = swin_blocks(x4, p + ".swin2", sd, heads, 2)
    x5 = patch_down(x4, p + ".down2", sd)
    x5 = swin_blocks(x5, p + ".swin3", sd, heads, 6)
    x5 = patch_up(x5, p + ".up2", sd)
    xx = x5 + x4
    xx = swin_blocks(xx, p + ".swin4", sd, heads, 2)
    xx = patch_up(xx, p + ".up1", sd)
    if scale_factor in (4, 8):
        xx = xx + F.linear(x3, sd[p + ".proj2.weight"], sd[p + ".proj2.bias"])
    else:
        xx = xx + x3
    xx = swin_blocks(xx, p + ".swin5", sd, heads, 2)
    return to_image(xx, p + ".to_image", sd, scale_factor)


def swin_unet_forward(sd, x, scale_factor, downscale_factor=1):
    """Eval forward of SwinUNet / SwinUNet2x / SwinUNet4x (swin_unet.py:221-226,
    :246-251, :280-287) and, with ``downscale_factor`` in {2,4},
    SwinUNetDownscaled (swin_unet.py:366-379: clamp, bicubic-AA resize, clamp)."""
    z = torch.clamp(swin_unet_base(sd, x, scale_factor), 0., 1.)
    if downscale_factor > 1:
        z = F.interpolate(z.float(), size=(z.shape[-2] // downscale_factor, z.shape[-1] // downscale_factor),
                          mode="bicubic", align_corners=False, antialias=True)
        z = torch.clamp(z, 0., 1.)
    return z


def bicubic_aa_weights(in_size, out_size):
    """Separable weights of ATen upsample_bicubic2d_aa (align_corners=False,
    A=-0.5): returns (start[out], w[out, taps]).  Restated from the documented
    ATen algorithm (aten/src/ATen/native/cpu/UpSampleKernel.cpp,
    ``HelperInterpBase::_compute_indices_min_size_weights_aa``) so the CUDA
    2x/4x downscale epilogue has an independent check."""
    import math
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    taps = int(math.ceil(support)) * 2 + 1

    def cubic(x, a=-0.5):
        x = abs(x)
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
        if x < 2.0:
            return (((x - 5.0) * x + 8.0) * x - 4.0) * a
        return 0.0

    starts, ws = [], []
    for i in range(out_size):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(int(center + support + 0.5), in_size) - xmin
        w = [cubic((j + xmin - center + 0.5) * invscale) for j in range(xsize)]
        tot = sum(w)
        w = [v / tot for v in w] + [0.0] * (taps - xsize)
        starts.append(xmin)
        ws.append(w)
    return torch.tensor(starts), torch.tensor(ws, dtype=torch.float32)


SWIN = {
    # name: (unet scale_factor, downscale, i2i scale, offset, blend)  swin_unet.py:213,234,267,345-350
    "swin_unet_1x": (1, 1, 1, 8, 4),
    "swin_unet_2x": (2, 1, 2, 16, 8),
    "swin_unet_4x": (4, 1, 4, 32, 16),
    "swin_unet_4x_to_2x": (4, 2, 2, 16, 8),
    "swin_unet_4x_to_1x": (4, 4, 1, 8, 16),
}
