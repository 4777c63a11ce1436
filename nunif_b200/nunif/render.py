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


def tiled_render(x, model, tile_size=None, batch_size=None, enable_amp=True, out=None, non_blocking=False):
    """x: C,H,W float tensor -> C,H*scale,W*scale, contiguous, clamped.

    * x on the model's device: the result is a device tensor (render.py:8-19).
    * x on the HOST (the reference accepts that too - SeamBlending moves each minibatch with ``.to(device)``,
      seam_blending.py:94): ``nb200_tiled_render_host`` copies the frame in once and streams the blended output back
      in bands of finished tile rows while later tile batches compute.  The result is a pinned host tensor
      (``out`` if given: C,H*scale,W*scale float32, ideally pinned).  Like the reference, the call returns a
      COMPLETED tensor; ``non_blocking=True`` (with a pinned ``x``) skips the final stream synchronisation for
      callers that pipeline frames - the result is then complete when the model device's current stream is.

    The engine implements the reference's CUDA numerics (fp16 autocast, nunif/device.py:58-71) and nothing else:
    ``enable_amp=False`` (an fp32 forward) raises instead of silently running fp16.
    """
    if not enable_amp:
        raise NotImplementedError("nunif_b200 runs the reference's CUDA autocast (fp16) numerics only; "
                                  "enable_amp=False (fp32 forward) is not implemented")
    assert not torch.is_grad_enabled()                                # seam_blending.py:50
    assert x.ndim == 3 and x.shape[0] == 3
    batch_size = batch_size or model.i2i_default_batch_size
    tile_size = model.find_valid_tile_size(tile_size)
    C, H, W = x.shape
    oshape = (C, H * model.i2i_scale, W * model.i2i_scale)
    if not x.is_cuda:
        dev = model.device
        xf = x.float().contiguous()
        if out is None:
            out = torch.empty(oshape, dtype=torch.float32, pin_memory=True)
        assert (not out.is_cuda) and out.dtype == torch.float32 and tuple(out.shape) == oshape and out.is_contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nb200_tiled_render_host(model._h, _lib.ptr(xf), C, H, W, int(tile_size), int(batch_size),
                                                          int(model._downscale), _lib.ptr(out), _lib.stream_ptr(dev)))
            if not (non_blocking and xf.is_pinned()):
                torch.cuda.current_stream(dev).synchronize()           # band D2H copies have landed; xf may be a temporary
        return out
    _lib.require_cuda(x, "x")
    if x.device != model.device:
        x = x.to(model.device)                                         # seam_blending.py:94 moves each minibatch to the model
    xf = x.float().contiguous()
    if out is None:
        out = torch.empty(oshape, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().nb200_tiled_render(model._h, _lib.ptr(xf), C, H, W, int(tile_size), int(batch_size),
                                                 int(model._downscale), _lib.ptr(out), _lib.stream_ptr(x.device)))
    return out.to(x.dtype) if x.dtype == torch.float16 else out
