"""Mirror of iw3/zoedepth_model.py:30-85 (batch_preprocess): size rule + fused antialiased resize /
reflection pad / clamp / normalise (csrc/frame_ops.cu).  The ZoeDepth network body is not part of this engine yet."""
import ctypes
import torch
from .. import _lib


def preprocess_size(H, W, h_height=384, v_height=512, ensure_multiple_of=32):
    """-> (new_h, new_w, pad_h, pad_w, frame_h, frame_w); host integers, no GPU needed."""
    v = [ctypes.c_int(0) for _ in range(6)]
    _lib.check(_lib.lib().nb200_zoe_preprocess_size(int(H), int(W), int(h_height), int(v_height), int(ensure_multiple_of),
                                                    *[ctypes.byref(a) for a in v]))
    return tuple(a.value for a in v)


def batch_preprocess(x, h_height=384, v_height=512, ensure_multiple_of=32):
    """x: BCHW float32 0-1 (CUDA) -> (normalised, reflection-padded network input, pad_h, pad_w)."""
    _lib.require_cuda(x, "x")
    assert x.ndim == 4 and x.shape[1] == 3
    B, _, H, W = x.shape
    nh, nw, ph, pw, fh, fw = preprocess_size(H, W, h_height, v_height, ensure_multiple_of)
    xf = x.float().contiguous()
    out = torch.empty((B, 3, fh + 2 * ph, fw + 2 * pw), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().nb200_zoe_preprocess(_lib.ptr(xf), B, H, W, fh, fw, ph, pw, _lib.ptr(out), _lib.stream_ptr(x.device)))
    return out, ph, pw
