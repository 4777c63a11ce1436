"""Mirror of iw3/anaglyph.py:95-110 (apply_anaglyph_redcyan, all red-cyan methods)."""
import torch
from .. import _lib
from ._common import prep

TYPES = {"dubois": 0, "dubois2": 1, "color": 2, "gray": 3, "half-color": 4, "wimmer": 5, "wimmer2": 6}


def apply_anaglyph_redcyan(left_eye, right_eye, anaglyph_type):
    if anaglyph_type not in TYPES:
        raise ValueError(f"Unknown anaglyph_type {anaglyph_type}")      # anaglyph.py:110
    squeeze = left_eye.ndim == 3
    l = prep(left_eye.unsqueeze(0) if squeeze else left_eye, "left_eye")
    r = prep(right_eye.unsqueeze(0) if squeeze else right_eye, "right_eye")
    B, _, H, W = l.shape
    out = torch.empty_like(l)
    with torch.cuda.device(l.device):
        _lib.check(_lib.lib().nb200_anaglyph(_lib.ptr(l), _lib.ptr(r), B, H, W, TYPES[anaglyph_type], _lib.ptr(out),
                                             _lib.stream_ptr(l.device)))
    return out[0] if squeeze else out
