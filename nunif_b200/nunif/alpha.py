"""Mirror of nunif/utils/alpha.py:32-57 (AlphaBorderPadding)."""
import torch
from .. import _lib


class AlphaBorderPadding:
    """rgb 3,H,W and alpha 1,H,W (CUDA, float) -> rgb with transparent pixels filled from opaque neighbours.

    ``offset`` rounds of a fused 3x3 kernel (csrc/alpha_tta.cu); callable like the reference nn.Module."""

    def eval(self):
        return self

    def to(self, device):
        return self

    def __call__(self, rgb, alpha, offset):
        return self.forward(rgb, alpha, offset)

    def forward(self, rgb, alpha, offset):
        assert rgb.ndim == 3 and alpha.ndim == 3 and rgb.shape[0] == 3 and alpha.shape[0] == 1   # alpha.py:41
        _lib.require_cuda(rgb, "rgb")
        _lib.require_cuda(alpha, "alpha")
        rgbf, af = rgb.float().contiguous(), alpha.float().contiguous()
        _, H, W = rgbf.shape
        out = torch.empty_like(rgbf)
        ws = torch.empty(_lib.lib().nb200_alpha_border_padding_workspace(H, W), dtype=torch.uint8, device=rgb.device)
        with torch.cuda.device(rgb.device):
            _lib.check(_lib.lib().nb200_alpha_border_padding(_lib.ptr(rgbf), _lib.ptr(af), H, W, int(offset), _lib.ptr(out),
                                                             _lib.ptr(ws), _lib.stream_ptr(rgb.device)))
        return out.to(rgb.dtype)
