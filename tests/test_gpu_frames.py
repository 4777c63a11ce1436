"""GPU parity of the frame-edge conversions and DepthAnything's batch_preprocess."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, t, log_metric, stats
from nunif_b200 import synth
from oracle import frames as ofr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_hwc_to_chw_float_exact():
    from nunif_b200.iw3 import hwc_to_chw_float
    g = load_golden("frames")
    assert torch.equal(hwc_to_chw_float(torch.from_numpy(g["u8"]), DEV).cpu(), t(g["u8_f"]))
    u16 = torch.from_numpy(g["u16"]).view(torch.uint16)
    assert torch.equal(hwc_to_chw_float(u16, DEV).cpu(), t(g["u16_f"]))
    # 1080p batch against the oracle
    x = torch.randint(0, 256, (3, 1080, 1920, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    assert np.array_equal(hwc_to_chw_float(x, DEV).cpu().numpy(), ofr.hwc_to_chw_float(x.numpy()))
    # float input is only permuted (iw3/utils.py:277-282)
    f = torch.rand(5, 7, 3)
    assert torch.equal(hwc_to_chw_float(f, DEV).cpu(), f.permute(2, 0, 1))


def test_chw_float_to_hwc_exact_round_half_even():
    from nunif_b200.iw3 import chw_float_to_hwc, hwc_to_chw_float
    g = load_golden("frames")
    f = t(g["f"], DEV)
    assert torch.equal(chw_float_to_hwc(f).cpu(), torch.from_numpy(g["f_u8"]))
    assert torch.equal(chw_float_to_hwc(f, use_16bit=True).cpu().view(torch.int16), torch.from_numpy(g["f_u16"]))
    # uint8 -> float -> uint8 is the identity at 1080p (round trip property)
    x = torch.randint(0, 256, (2, 1080, 1920, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(2))
    assert torch.equal(chw_float_to_hwc(hwc_to_chw_float(x, DEV)).cpu(), x)


def test_da_batch_preprocess_golden():
    from nunif_b200.iw3 import batch_preprocess
    g = load_golden("frames")
    for key, src, kw in (("prep_126", "x", dict(lower_bound=126)), ("prep_98_limit", "x", dict(lower_bound=392, limit_resolution=True)),
                         ("prep_tall", "xt", dict(lower_bound=70))):
        got = batch_preprocess(t(g[src], DEV), **kw)
        s = stats(got, t(g[key]))
        log_metric("da_" + key, **s)
        assert got.shape == g[key].shape and s["max"] < 5e-6, (key, s)     # values are O(1)..O(2.6) after normalisation


def test_da_batch_preprocess_1080p_against_oracle():
    from nunif_b200.iw3 import batch_preprocess
    x = synth.synth_image(8, 3, 1080, 1920, smooth=False).unsqueeze(0)
    got = batch_preprocess(x.to(DEV))
    assert got.shape == (1, 3, 392, 686)
    want = ofr.batch_preprocess(x.numpy())
    s = stats(got, torch.from_numpy(want))
    log_metric("da_prep_1080p", **s)
    assert s["max"] < 5e-6


def test_zoe_batch_preprocess_golden():
    from nunif_b200.iw3.zoedepth_preprocess import batch_preprocess
    g = load_golden("frames")
    for key, src in (("zoe_land", "x"), ("zoe_port", "xt")):
        got, ph, pw = batch_preprocess(t(g[src], DEV), h_height=96, v_height=128)
        assert (ph, pw) == tuple(g[key + "_pad"]) and got.shape == g[key].shape
        s = stats(got, t(g[key]))
        log_metric(key, **s)
        assert s["max"] < 5e-6, (key, s)
    # 4K frame (BASELINE configs[4] geometry: 384x704 with pads 16, 22) against the oracle
    x = synth.synth_image(12, 3, 2160, 3840, smooth=False).unsqueeze(0)
    got, ph, pw = batch_preprocess(x.to(DEV))
    assert got.shape == (1, 3, 384, 704) and (ph, pw) == (16, 22)
    want, _, _ = ofr.zoe_batch_preprocess(x.numpy())
    assert stats(got, torch.from_numpy(want))["max"] < 5e-6
