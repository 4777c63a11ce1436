"""Device-side frame conversions at the edge of the iw3 path.

Mirrors nunif/utils/video.py:218-223 (to_tensor), :236-246 (from_tensor) and
iw3/utils.py:274-289 (hwc_to_chw_float); csrc/frame_ops.cu."""
import torch
from .. import _lib

_BITS = {torch.uint8: 8, torch.uint16: 16}


def hwc_to_chw_float(x, device=None):
    """uint8/uint16 HWC or BHWC (any device) -> float32 CHW / BCHW in [0,1] on ``device`` (iw3/utils.py:274-289).
    Floating-point inputs are only permuted, as in the reference."""
    if device is not None:
        x = x.to(device)
    _lib.require_cuda(x, "x")
    if x.ndim not in (3, 4):
        raise ValueError(f"Unsupported ndim={x.ndim}")
    if torch.is_floating_point(x):
        return (x.permute(2, 0, 1) if x.ndim == 3 else x.permute(0, 3, 1, 2)).contiguous()
    if x.dtype not in _BITS or x.shape[-1] != 3:
        raise ValueError("expected uint8/uint16 frames with 3 interleaved channels")
    xc = x.contiguous()
    B = 1 if x.ndim == 3 else x.shape[0]
    H, W = x.shape[-3], x.shape[-2]
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().nb200_hwc_to_chw_f32(_lib.ptr(xc), _BITS[x.dtype], B, H, W, _lib.ptr(out), _lib.stream_ptr(x.device)))
    return out[0] if x.ndim == 3 else out


def chw_float_to_hwc(x, use_16bit=False):
    """float CHW / BCHW in [0,1] (CUDA) -> uint8 (or uint16) HWC / BHWC on the device:
    ``(x.permute(1, 2, 0) * scale).round_().to(dtype)`` of from_tensor (video.py:236-246)."""
    _lib.require_cuda(x, "x")
    if x.ndim not in (3, 4) or x.shape[-3] != 3:
        raise ValueError("expected CHW or BCHW with 3 channels")
    xf = x.float().contiguous()
    B = 1 if x.ndim == 3 else x.shape[0]
    H, W = x.shape[-2], x.shape[-1]
    dtype = torch.uint16 if use_16bit else torch.uint8
    out = torch.empty((B, H, W, 3), dtype=dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().nb200_chw_f32_to_hwc(_lib.ptr(xf), 16 if use_16bit else 8, B, H, W, _lib.ptr(out), _lib.stream_ptr(x.device)))
    return out[0] if x.ndim == 3 else out
