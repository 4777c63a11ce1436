"""GPU parity of iw3's output composition: all anaglyph methods, the bicubic-antialias resize and
postprocess_image against outputs of the reference's own functions (tests/golden/anaglyph.npz, postprocess.npz)."""
import types
import numpy as np
import pytest
import torch

from tests.util import load_golden, t, log_metric, stats
from tests.test_oracle_golden import _postprocess_cases
from nunif_b200 import synth
from oracle import postprocess as opp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_anaglyph_all_methods_golden():
    from nunif_b200.iw3 import apply_anaglyph_redcyan
    g = load_golden("anaglyph")
    l, r = t(g["l"], DEV), t(g["r"], DEV)
    for kind in ("color", "gray", "half-color", "wimmer", "wimmer2", "dubois", "dubois2"):
        s = stats(apply_anaglyph_redcyan(l, r, kind), t(g[kind.replace("-", "_")]))
        log_metric("anaglyph_" + kind, **s)
        assert s["max"] < (2e-5 if kind.startswith("dubois") or kind == "wimmer2" else 1e-7), (kind, s)
    with pytest.raises(ValueError):
        apply_anaglyph_redcyan(l, r, "nope")


def test_postprocess_image_golden():
    from nunif_b200.iw3 import postprocess_image
    g = load_golden("postprocess")
    l, r = t(g["l"], DEV), t(g["r"], DEV)
    for name, kw in _postprocess_cases():
        got = postprocess_image(l, r, types.SimpleNamespace(**kw))
        want = t(g["pp_" + name])
        assert tuple(got.shape) == tuple(want.shape), (name, got.shape, want.shape)
        s = stats(got, want)
        log_metric("postprocess_" + name, **s)
        assert s["max"] < 2e-5, (name, s)


def test_vr180_equirectangular_golden_and_oracle():
    """iw3/equirectangular.py through the reference-generated golden (tests/golden/postprocess.npz vr180_l) and, for a
    non-square odd-sized frame, the numpy oracle; then the vr180 branch of postprocess_image."""
    from nunif_b200.iw3 import equirectangular_projection, postprocess_image
    g = load_golden("postprocess")
    l, r = t(g["l"], DEV).clamp(0, 1), t(g["r"], DEV).clamp(0, 1)
    got = equirectangular_projection(l)
    want = t(g["vr180_l"])
    assert tuple(got.shape) == tuple(want.shape)
    s = stats(got, want)
    log_metric("vr180_golden", **s)
    assert s["max"] < 2e-4 and s["mean"] < 2e-6, s
    x = synth.synth_image(9, 3, 75, 133)
    got = equirectangular_projection(x.to(DEV))
    want = torch.from_numpy(opp.equirectangular_projection(x.numpy()))
    assert tuple(got.shape) == tuple(want.shape)
    s = stats(got, want)
    log_metric("vr180_oracle_odd", **s)
    assert s["max"] < 2e-4 and s["mean"] < 2e-6, s
    sbs = postprocess_image(l, r, types.SimpleNamespace(vr180=True))
    assert sbs.shape[1] == got.shape[1] * 0 + equirectangular_projection(l).shape[1] and sbs.shape[2] == 2 * equirectangular_projection(l).shape[2]
    assert torch.equal(sbs[:, :, :sbs.shape[2] // 2], equirectangular_projection(l))


def test_half_sbs_1080p_against_oracle():
    from nunif_b200.iw3 import resize_bicubic_aa
    x = synth.synth_image(41, 3, 1080, 1920, smooth=False)
    got = resize_bicubic_aa(x.to(DEV), (1080, 960))
    want = opp.resize_bicubic_aa(x.numpy(), 1080, 960)
    assert stats(got, torch.from_numpy(want))["max"] < 1e-5
    # identity when the size does not change (scale 1: the cubic kernel is 1 at 0 and 0 at the other integers)
    assert stats(resize_bicubic_aa(x.to(DEV), (1080, 1920)), x)["max"] < 1e-6
