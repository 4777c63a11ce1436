"""Mirror of iw3/depth_anything_model.py:69-110 (batch_preprocess): resize rule + fused
antialiased resize / clamp / ImageNet normalise (csrc/frame_ops.cu)."""
import ctypes
import torch
from .. import _lib


def preprocess_size(H, W, lower_bound=392, max_aspect_ratio=4, limit_resolution=False):
    """(new_h, new_w) of batch_preprocess (depth_anything_model.py:69-101); host integers, no GPU needed."""
    nh, nw = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.lib().nb200_da_preprocess_size(int(H), int(W), int(lower_bound), int(max_aspect_ratio),
                                                   1 if limit_resolution else 0, ctypes.byref(nh), ctypes.byref(nw)))
    return nh.value, nw.value


def batch_preprocess(x, lower_bound=392, max_aspect_ratio=4, limit_resolution=False):
    """x: BCHW float32 0-1 (CUDA) -> B,3,new_h,new_w normalised network input."""
    _lib.require_cuda(x, "x")
    assert x.ndim == 4 and x.shape[1] == 3
    B, _, H, W = x.shape
    nh, nw = preprocess_size(H, W, lower_bound, max_aspect_ratio, limit_resolution)
    xf = x.float().contiguous()
    out = torch.empty((B, 3, nh, nw), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().nb200_da_preprocess(_lib.ptr(xf), B, H, W, nh, nw, _lib.ptr(out), _lib.stream_ptr(x.device)))
    return out
