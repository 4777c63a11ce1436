"""Mirror of the per-frame path of iw3/depth_scaler.py:4-17 + base_depth_model.py:176-194,
with the disparity mapper (iw3/mapper.py) fused into the same pass."""
import torch
from .. import _lib
from ._common import prep

_DIV_C = {"div_25": 2.5, "div_10": 1.0, "div_6": 0.6, "div_4": 0.4, "div_2": 0.2, "div_1": 0.1}  # mapper.py:106-113


def minmax_normalize(depth, mapper="none", return_minmax=False):
    """depth: B,1,h,w (or 1,h,w): per-frame (x-min)/(max-min) clamp[0,1], then mapper
    ("none" or "div_*").  No host sync (the reference's ``if scale > 0`` syncs)."""
    squeeze = depth.ndim == 3
    d = prep(depth.unsqueeze(0) if squeeze else depth, "depth")
    if mapper == "none":
        c = -1.0
    elif mapper in _DIV_C:
        c = _DIV_C[mapper]
    else:
        raise NotImplementedError(f"mapper={mapper}")
    B = d.shape[0]
    n = d[0].numel()
    out = torch.empty_like(d)
    mm = torch.empty((B, 2), device=d.device, dtype=torch.float32) if return_minmax else None
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().nb200_minmax_map(_lib.ptr(d), B, n, c, _lib.ptr(out), _lib.ptr(mm), _lib.stream_ptr(d.device)))
    out = out[0] if squeeze else out
    return (out, mm) if return_minmax else out
