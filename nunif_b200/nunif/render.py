"""Mirror of nunif/utils/render.py:8-19 (tiled_render) and
nunif/utils/seam_blending.py:109-143 (create_config)."""
import ctypes
import torch
from .. import _lib


def create_config(x_size, scale, offset, tile_size, blend_size):
    """SeamBlending.create_config - computed by the library's integer planner."""
    cfg = _lib.TileConfig()
    _lib.check(_lib.lib().nb200_tile_config_create(int(x_size[0]), int(x_size[1]), int(scale), int(offset),
                                                   int(tile_size), int(blend_size), ctypes.byref(cfg)))
    return cfg.as_dict()


def tiled_render(x, model, tile_size=None, batch_size=None, enable_amp=False):
    """x: C,H,W float tensor on the model's device -> C,H*scale,W*scale, contiguous, clamped.

    The engine always runs the reference's CUDA numerics (fp16 autocast,
    nunif/device.py:58-71), so ``enable_amp`` is accepted for signature parity only.
    """
    assert not torch.is_grad_enabled()                                # seam_blending.py:50
    _lib.require_cuda(x, "x")
    assert x.ndim == 3 and x.shape[0] == 3
    xf = x.float().contiguous()
    C, H, W = xf.shape
    batch_size = batch_size or model.i2i_default_batch_size
    tile_size = model.find_valid_tile_size(tile_size)
    out = torch.empty((C, H * model.i2i_scale, W * model.i2i_scale), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().nb200_tiled_render(model._h, _lib.ptr(xf), C, H, W, int(tile_size), int(batch_size),
                                                 int(model._downscale), _lib.ptr(out), _lib.stream_ptr(x.device)))
    return out.to(x.dtype) if x.dtype == torch.float16 else out
