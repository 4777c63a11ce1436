import torch
from .. import _lib

VIEWS = {"both": 0, "left": 1, "right": 2}
COMPOSE_NONE, COMPOSE_SBS, COMPOSE_ANAGLYPH = 0, 1, 2


def prep(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
