"""Mirror of iw3's learned stereo warp: the `sbs.row_flow_v3` model (iw3/models/row_flow_v3.py) in delta_output mode and
its driver apply_divergence_nn_LR / apply_divergence_nn_delta (iw3/backward_warp.py:124-232).

The delta network runs as tcgen05 GEMMs + the kernels in csrc/rowflow_kernels.cu; the warp is the fused grid-sample kernel
(csrc/warp_backward.cu, nb200_backward_warp_delta).  steps > 1 (iterative re-warping of the depth) and
preserve_screen_border are not implemented and raise NotImplementedError.
"""
import ctypes
import torch
from .. import _lib

KIND_ROW_FLOW_V3 = 7     # NB200_MODEL_ROW_FLOW_V3


class RowFlowV3:
    """Packed `sbs.row_flow_v3`; ``model(x)`` with x = B,3,h,w (depth, divergence feature, convergence feature) returns the
    delta B,2,h,w (x component, zero y component) like the reference with ``delta_output=True`` (row_flow_v3.py:111-116)."""
    name = "sbs.row_flow_v3"
    symmetric = False
    delta_output = True

    def __init__(self, state_dict, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("nunif_b200 models live on a CUDA (sm_100) device; there is no CPU path")
        items = [(k, v.detach().to("cpu", torch.float32).contiguous()) for k, v in state_dict.items()]
        n = len(items)
        names = (ctypes.c_char_p * n)(*[k.encode() for k, _ in items])
        datas = (ctypes.c_void_p * n)(*[v.data_ptr() for _, v in items])
        numels = (ctypes.c_int64 * n)(*[v.numel() for _, v in items])
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nb200_model_create(KIND_ROW_FLOW_V3, n, names, datas, numels, 0, ctypes.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if self._h:
                _lib.lib().nb200_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def delta_x(self, x):
        _lib.require_cuda(x, "x")
        assert x.ndim == 4 and x.shape[1] == 3
        B, _, h, w = x.shape
        xf = x.float().contiguous()
        out = torch.empty((B, 1, h, w), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().nb200_row_flow_delta(self._h, _lib.ptr(xf), B, h, w, _lib.ptr(out), _lib.stream_ptr(x.device)))
        return out

    def __call__(self, x):
        d = self.delta_x(x)
        return torch.cat([d, torch.zeros_like(d)], dim=1)


def make_divergence_feature_value(divergence, convergence, image_width):
    """iw3/backward_warp.py:8-15."""
    divergence_pix = divergence * 0.5 * 0.01 * image_width
    return divergence_pix / 32.0, (-divergence_pix * convergence) / 32.0


def _warp_delta(c, delta, delta_scale):
    B, _, H, W = c.shape
    h, w = delta.shape[-2:]
    out = torch.empty_like(c)
    with torch.cuda.device(c.device):
        _lib.check(_lib.lib().nb200_backward_warp_delta(_lib.ptr(c), _lib.ptr(delta), B, H, W, h, w, float(delta_scale), _lib.ptr(out),
                                                        _lib.stream_ptr(c.device)))
    return out


def apply_divergence_nn_delta(model, c, depth, divergence, convergence, steps, shift, preserve_screen_border=False, enable_amp=True):
    """iw3/backward_warp.py:185-232 (steps == 1)."""
    if steps not in (None, 1):
        raise NotImplementedError("steps > 1 is not implemented by the B200 engine")
    if preserve_screen_border:
        raise NotImplementedError("preserve_screen_border is not implemented by the B200 engine")
    _lib.require_cuda(c, "c")
    _lib.require_cuda(depth, "depth")
    c, depth = c.float().contiguous(), depth.float().contiguous()
    if shift > 0:
        c, depth = torch.flip(c, (3,)), torch.flip(depth, (3,))
    B, _, H, W = depth.shape
    dv, cv = make_divergence_feature_value(divergence, convergence, max(H, W))
    x = torch.cat([depth, torch.full_like(depth, dv), torch.full_like(depth, cv)], dim=1)      # make_input_tensor(None, ...)
    delta = model.delta_x(x)
    z = _warp_delta(c, delta, 1.0 / (W // 2 - 1))                                              # :201
    return torch.flip(z, (3,)) if shift > 0 else z


def apply_divergence_nn_LR(model, c, depth, divergence, convergence, steps=None, synthetic_view="both",
                           preserve_screen_border=False, enable_amp=True):
    """iw3/backward_warp.py:124-160 for the non-symmetric delta models."""
    assert synthetic_view in {"both", "right", "left"}
    if getattr(model, "symmetric", False) or getattr(model, "name", "") == "sbs.mlbw":
        raise NotImplementedError("only sbs.row_flow_v3 is implemented by the B200 engine")
    kw = dict(steps=steps, preserve_screen_border=preserve_screen_border, enable_amp=enable_amp)
    if synthetic_view == "both":
        return (apply_divergence_nn_delta(model, c, depth, divergence, convergence, shift=-1, **kw),
                apply_divergence_nn_delta(model, c, depth, divergence, convergence, shift=1, **kw))
    if synthetic_view == "right":
        return c, apply_divergence_nn_delta(model, c, depth, divergence * 2, convergence, shift=1, **kw)
    return apply_divergence_nn_delta(model, c, depth, divergence * 2, convergence, shift=-1, **kw), c
