"""GPU parity of AlphaBorderPadding / tta_split / tta_merge and of the
alpha + TTA branches of Waifu2x.convert (A2) through the C ABI."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, t, log_metric, stats
from nunif_b200 import synth
from oracle import alpha_tta as oat

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_alpha_border_padding_golden():
    from nunif_b200.nunif.alpha import AlphaBorderPadding
    g = load_golden("alpha_tta")
    pad = AlphaBorderPadding()
    for off in (0, 1, 8, 17, 36):
        got = pad(t(g["rgb"], DEV), t(g["alpha"], DEV), off)
        s = stats(got, t(g[f"pad_{off}"]))
        log_metric(f"alpha_pad_{off}", **s)
        assert s["max"] < 2e-6, (off, s)      # 3x3 box-sum order differs from the reference's depthwise conv


def test_alpha_border_padding_4k_against_oracle_band_and_properties():
    from nunif_b200.nunif.alpha import AlphaBorderPadding
    H, W = 2160, 3840
    rgb = synth.synth_image(11, 3, H, W)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    alpha = (((yy - 1000) ** 2 + (xx - 1900) ** 2).sqrt() < 700).float().unsqueeze(0)
    got = AlphaBorderPadding()(rgb.to(DEV), alpha.to(DEV), 32).cpu()
    opaque = alpha[0] > 0
    assert torch.equal(got[:, opaque], rgb[:, opaque])                     # opaque pixels are untouched
    far = ((yy - 1000) ** 2 + (xx - 1900) ** 2).sqrt() > 700 + 32 * 1.5    # beyond the reach of 32 one-pixel rounds
    assert float(got[:, far].abs().max()) == 0.0
    # a band around the circle against the oracle
    ys, xs = slice(250, 420), slice(1700, 2100)
    want = oat.alpha_border_padding(rgb[:, 200:470, 1650:2150].numpy(), alpha[:, 200:470, 1650:2150].numpy(), 32)
    assert np.abs(got[:, ys, xs].numpy() - want[:, 50:220, 50:450]).max() < 2e-6


def test_tta_split_merge_golden_exact():
    from nunif_b200.nunif.tta import tta_split, tta_merge
    g = load_golden("alpha_tta")
    views = tta_split(t(g["x"], DEV))
    for k in range(8):
        assert torch.equal(views[k].cpu(), t(g[f"view_{k}"])), k
    merged = tta_merge([t(g[f"z_{k}"], DEV) for k in range(8)])
    assert torch.equal(merged.cpu(), t(g["merged"]))
    ident = tta_merge(list(views))
    assert stats(ident, t(g["merged_identity"]))["max"] == 0.0
    # full-size round trip: merge(split(x)) == x up to the 1/8 averaging rounding
    x = synth.synth_image(5, 3, 1080, 1920).to(DEV)
    assert stats(tta_merge(list(tta_split(x))), x)["max"] < 1e-6


def test_convert_alpha_and_tta_branches(tmp_path):
    """Waifu2x.convert with a non-blank alpha and tta=True equals the same composition of the verified parts."""
    from nunif_b200.waifu2x.utils import Waifu2x
    from nunif_b200.nunif.models import create_model
    from nunif_b200.nunif.render import tiled_render
    from nunif_b200.nunif.alpha import AlphaBorderPadding
    from nunif_b200.nunif.tta import tta_split, tta_merge
    w = Waifu2x(str(tmp_path), [0])
    w.scale_model = create_model("waifu2x.upcunet", synth.upcunet_state_dict(0), DEV)
    x = synth.synth_image(9, 3, 96, 128)
    alpha = torch.ones(1, 96, 128)
    alpha[:, 20:60, 30:90] = 0
    with torch.inference_mode():
        rgb, a = w.convert(x, alpha, "scale", -1, tile_size=64, batch_size=4, tta=True)
        assert not rgb.is_cuda and rgb.shape == (3, 192, 256) and a.shape == (1, 192, 256)
        xp = AlphaBorderPadding()(x.to(DEV), alpha.to(DEV), w.scale_model.i2i_offset)
        want = tta_merge([tiled_render(v, w.scale_model, tile_size=64, batch_size=4) for v in tta_split(xp)])
        assert torch.equal(rgb, want.cpu())
        wa = tiled_render(alpha.to(DEV).expand(3, 96, 128), w.scale_model, tile_size=64, batch_size=4).mean(0, keepdim=True)
        assert torch.equal(a, wa.cpu())
        rgb2, a2 = w.convert(x, torch.ones(1, 96, 128), "scale", -1, tile_size=64, batch_size=4)
        assert torch.equal(a2, torch.ones(1, 192, 256))
        assert torch.equal(rgb2, tiled_render(x.to(DEV), w.scale_model, tile_size=64, batch_size=4).cpu())
