"""Mirror of iw3/anaglyph.py:95-110 for the dubois family (the hot-path default,
iw3/utils.py:2076)."""
import torch
from .. import _lib
from ._common import prep


def apply_anaglyph_redcyan(left_eye, right_eye, anaglyph_type):
    if anaglyph_type not in {"dubois", "dubois2"}:
        if anaglyph_type in {"color", "gray", "half-color", "wimmer", "wimmer2"}:
            raise NotImplementedError(f"anaglyph_type {anaglyph_type} is outside the B200 hot path")
        raise ValueError(f"Unknown anaglyph_type {anaglyph_type}")      # anaglyph.py:110
    squeeze = left_eye.ndim == 3
    l = prep(left_eye.unsqueeze(0) if squeeze else left_eye, "left_eye")
    r = prep(right_eye.unsqueeze(0) if squeeze else right_eye, "right_eye")
    B, _, H, W = l.shape
    out = torch.empty_like(l)
    with torch.cuda.device(l.device):
        _lib.check(_lib.lib().nb200_anaglyph_dubois(_lib.ptr(l), _lib.ptr(r), B, H, W,
                                                    1 if anaglyph_type == "dubois" else 0,
                                                    _lib.ptr(out), _lib.stream_ptr(l.device)))
    return out[0] if squeeze else out
