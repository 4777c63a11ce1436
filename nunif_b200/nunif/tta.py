"""Mirror of nunif/transforms/tta.py:20-48 (tta_split / tta_merge) on the device."""
import ctypes
import torch
from .. import _lib


def tta_split(x):
    """x C,H,W (CUDA float) -> the 8 dihedral views in the reference's order (tta.py:20-33)."""
    assert (isinstance(x, torch.Tensor) and x.dim() == 3)
    _lib.require_cuda(x, "x")
    xf = x.float().contiguous()
    C, H, W = xf.shape
    views = []
    with torch.cuda.device(x.device):
        for k in range(8):
            out = torch.empty((C, H, W) if k < 4 else (C, W, H), dtype=torch.float32, device=x.device)
            _lib.check(_lib.lib().nb200_tta_transform(_lib.ptr(xf), C, H, W, k, _lib.ptr(out), _lib.stream_ptr(x.device)))
            views.append(out.to(x.dtype))
    return tuple(views)


def tta_merge(xs):
    """xs: the 8 processed views (same order) -> clamp(mean of the inverse-transformed views) (tta.py:36-48)."""
    assert len(xs) == 8
    for v in xs:
        _lib.require_cuda(v, "view")
    vs = [v.float().contiguous() for v in xs]
    C, H, W = vs[0].shape
    for k, v in enumerate(vs):
        assert tuple(v.shape) == ((C, H, W) if k < 4 else (C, W, H)), "view shapes do not form a dihedral set"
    out = torch.empty((C, H, W), dtype=torch.float32, device=vs[0].device)
    ptrs = (ctypes.c_void_p * 8)(*[v.data_ptr() for v in vs])
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().nb200_tta_merge(ptrs, C, H, W, _lib.ptr(out), _lib.stream_ptr(out.device)))
    return out.to(xs[0].dtype)
