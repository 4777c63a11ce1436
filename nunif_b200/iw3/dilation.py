"""Mirror of iw3/dilation.py:5-27,115-142."""
import torch
from .. import _lib
from ._common import prep


def edge_dilation_parse(edge_dilation):
    """dilation.py:5-22 (same accepted types and ValueError)."""
    if isinstance(edge_dilation, (list, tuple)):
        if len(edge_dilation) == 0:
            x = y = 0
        elif len(edge_dilation) == 1:
            x = y = edge_dilation[0]
        else:
            x, y = edge_dilation[0], edge_dilation[1]
    elif isinstance(edge_dilation, int):
        x = y = edge_dilation
    elif edge_dilation is None:
        x = y = 0
    else:
        raise ValueError(f"Unsupported edge_dilation type {type(edge_dilation)}. "
                         "Supported types: int, list, tuple.")
    return x, y


def edge_dilation_is_enabled(edge_dilation):
    """dilation.py:25-27."""
    x, y = edge_dilation_parse(edge_dilation)
    return x != 0 or y != 0


@torch.inference_mode()
def dilate_edge(x, n):
    """x: B,1,h,w float CUDA tensor; n: int or [x_iter, y_iter] -> new tensor."""
    x_iter, y_iter = edge_dilation_parse(n)
    x = prep(x, "x")
    assert x.ndim == 4                                       # dilation.py:102
    B, _, h, w = x.shape
    out = torch.empty_like(x)
    lib = _lib.lib()
    ws = torch.empty(lib.nb200_dilate_edge_workspace(B, h, w), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.nb200_dilate_edge(_lib.ptr(x), B, h, w, int(x_iter), int(y_iter), _lib.ptr(out),
                                         _lib.ptr(ws), _lib.stream_ptr(x.device)))
    return out
