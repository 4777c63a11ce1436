"""CPU-only: the C-ABI library builds, loads without a GPU, exports every symbol that
include/nunif_b200.h declares, and fails loudly (no fallback) when asked to compute."""
import ctypes
import os
import re
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from nunif_b200 import build, _lib
    build.build()
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nunif_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from nunif_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nunif_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in nunif_b200/_lib.py"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but not declared in the header"


def test_abi_version(lib):
    assert lib.nb200_abi_version() == 1


def test_tile_config_host_planner_bit_exact(lib):
    """create_config is host integer code: check it against the reference goldens without a GPU."""
    from nunif_b200.nunif.render import create_config
    from tests.util import load_golden
    g = load_golden("seam_config")
    for case, want in zip(g["cases"], g["configs"]):
        h, w, scale, offset, tile, blend = (int(v) for v in case)
        p = create_config((h, w), scale, offset, tile, blend)
        got = [p["y_h"], p["y_w"], p["h_blocks"], p["w_blocks"], *p["pad"],
               p["y_buffer_h"], p["y_buffer_w"], p["input_tile_step"], p["output_tile_step"]]
        assert got == [int(v) for v in want], (case, got, want)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from nunif_b200.iw3 import apply_divergence_grid_sample, dilate_edge
    from nunif_b200.nunif.models import create_model
    from nunif_b200 import synth
    assert lib.nb200_check_device(0) != 0
    assert b"no CPU fallback" in lib.nb200_last_error()
    with pytest.raises(RuntimeError):
        apply_divergence_grid_sample(torch.rand(1, 3, 8, 8), torch.rand(1, 1, 8, 8), 2.0, 0.5)
    with pytest.raises(RuntimeError):
        dilate_edge(torch.rand(1, 1, 8, 8), 2)
    with pytest.raises(RuntimeError):
        create_model("waifu2x.upcunet", synth.upcunet_state_dict(0), device="cpu")


def test_find_valid_tile_size_host():
    from tests.util import load_golden
    from nunif_b200.nunif import models
    g = load_golden("tile_size")
    m_c = object.__new__(models.B200I2IModel)
    m_c._validator = models._cunet_validator
    m_c.i2i_default_tile_size = 256
    m_s = object.__new__(models.B200I2IModel)
    m_s._validator = models._swin_validator
    m_s.i2i_default_tile_size = 256
    for q, c, s in zip(g["query"], g["cunet"], g["swin"]):
        assert m_c.find_valid_tile_size(int(q)) == int(c)
        assert m_s.find_valid_tile_size(int(q)) == int(s)
    assert m_s.find_valid_tile_size(None) == 256


def test_da_preprocess_size_host_rule_bit_exact(lib):
    """nb200_da_preprocess_size is host integer logic (depth_anything_model.py:69-101): check it against the sizes the
    reference produced (tests/golden/frames.npz) without a GPU."""
    from tests.util import load_golden
    from nunif_b200.iw3.depth_anything_preprocess import preprocess_size
    g = load_golden("frames")
    for H, W, lb, lim, nh, nw in g["sizes"]:
        assert preprocess_size(int(H), int(W), int(lb), 4, bool(lim)) == (int(nh), int(nw)), (H, W, lb, lim)
    assert preprocess_size(1080, 1920) == (392, 686)


def test_zoe_preprocess_size_host_rule_bit_exact(lib):
    """nb200_zoe_preprocess_size (zoedepth_model.py:30-71, incl. Python's round-half-even) against the reference's sizes."""
    from tests.util import load_golden
    from nunif_b200.iw3.zoedepth_preprocess import preprocess_size
    g = load_golden("frames")
    for H, W, oh, ow, ph, pw in g["zoe_sizes"]:
        nh, nw, p_h, p_w, fh, fw = preprocess_size(int(H), int(W))
        assert (fh + 2 * p_h, fw + 2 * p_w, p_h, p_w) == (int(oh), int(ow), int(ph), int(pw)), (H, W)
