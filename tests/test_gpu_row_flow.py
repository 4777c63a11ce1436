"""GPU parity of iw3's default learned stereo warp, sbs.row_flow_v3 + apply_divergence_nn_LR,
against golden outputs of the real reference model (tests/golden/row_flow.npz) and the oracle's restatement."""
import pytest
import torch

from tests.util import load_golden, t, log_metric, stats
from nunif_b200 import synth
from oracle import row_flow as orf

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _amp_delta(sd, x):
    sdc = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return orf.row_flow_delta(sdc, x.to(DEV)).float().cpu()


def _oracle_lr_on_device(sdc, c, d, div, conv):
    """oracle.row_flow.apply_divergence_nn_LR with its CPU-built mesh moved to the device."""
    def one(shift):
        cc, dd = (torch.flip(c, (3,)), torch.flip(d, (3,))) if shift > 0 else (c, d)
        delta = orf.row_flow_delta(sdc, orf.make_input(dd, div, conv))
        z = orf.warp_delta(cc.cpu(), delta.cpu(), dd.shape[3])
        return torch.flip(z, (3,)) if shift > 0 else z
    return one(-1), one(1)


def test_row_flow_delta_golden():
    from nunif_b200.iw3 import RowFlowV3
    g = load_golden("row_flow")
    sd = synth.row_flow_v3_state_dict(0)
    m = RowFlowV3(sd, DEV)
    x = t(g["x"], DEV)
    got = m(x)
    assert got.shape == (2, 2, 70, 130) and float(got[:, 1].abs().max()) == 0.0
    ref32, refamp = t(g["delta"]), _amp_delta(sd, t(g["x"]))
    e_ref, e_our = stats(refamp, ref32), stats(got[:, :1], ref32)
    log_metric("row_flow_delta", ours_max=e_our["max"], ours_mean=e_our["mean"], refamp_max=e_ref["max"], refamp_mean=e_ref["mean"],
               scale=float(ref32.abs().max()))
    scale = float(ref32.abs().max())
    assert e_our["max"] <= max(1e-3 * scale, 1.5 * e_ref["max"]) and e_our["mean"] <= max(5e-4 * scale, 1.25 * e_ref["mean"]), (e_our, e_ref)


def test_backward_warp_delta_exact_formula():
    """The delta warp alone (fp32 kernel) against the oracle's grid_sample on the reference's own delta."""
    from nunif_b200.iw3.row_flow import _warp_delta
    g = load_golden("row_flow")
    c, delta = t(g["c"], DEV), t(g["delta"], DEV)
    got = _warp_delta(c, delta, 1.0 / (130 // 2 - 1))
    want = orf.warp_delta(t(g["c"]), t(g["delta"]), 130)
    assert stats(got, want)["max"] < 1e-4


@pytest.mark.parametrize("key,dk,ck,div,conv,sv", [("", "d", "c", 2.0, 0.5, "both"), ("sv_right_", "d", "c", 2.5, 0.3, "right"),
                                                    ("2", "d2", "c2", 4.0, 0.6, "both")])
def test_apply_divergence_nn_LR_golden(key, dk, ck, div, conv, sv):
    from nunif_b200.iw3 import RowFlowV3, apply_divergence_nn_LR
    g = load_golden("row_flow")
    sd = synth.row_flow_v3_state_dict(0)
    m = RowFlowV3(sd, DEV)
    c, d = t(g[ck], DEV), t(g[dk], DEV)
    l, r = apply_divergence_nn_LR(m, c, d, div, conv, steps=1, synthetic_view=sv)
    wl, wr = (t(g["left" + key]), t(g["right" + key])) if key in ("", "2") else (t(g["sv_right_l"]), t(g["sv_right_r"]))
    # the delta is computed in fp16 (reference numerics under autocast; the golden is the reference's fp32 CPU run):
    # a delta error of e pixels moves the sample by e/2 source pixels, i.e. |dz| <= e/2 * max gradient of c
    sl, sr = stats(l, wl), stats(r, wr)
    log_metric("row_flow_lr_" + (key or "both"), left_max=sl["max"], right_max=sr["max"], left_mean=sl["mean"], right_mean=sr["mean"])
    assert sl["mean"] < 1e-3 and sr["mean"] < 1e-3 and sl["max"] < 3e-2 and sr["max"] < 3e-2, (sl, sr)
    if sv == "right":
        assert torch.equal(l, c)


def test_row_flow_1080p_runs_and_matches_oracle_amp():
    from nunif_b200.iw3 import RowFlowV3, apply_divergence_nn_LR
    sd = synth.row_flow_v3_state_dict(1)
    m = RowFlowV3(sd, DEV)
    c = synth.synth_image(31, 3, 1080, 1920).unsqueeze(0).to(DEV)
    d = synth.synth_depth(32, 1, 392, 686).to(DEV)
    l, r = apply_divergence_nn_LR(m, c, d, 2.0, 0.5)
    # the fp32 oracle evaluated on the GPU by torch (a 1080p frame through it takes minutes on the CPU)
    sdc = {k: v.to(DEV) for k, v in sd.items()}
    from tests.util import true_fp32
    with torch.no_grad(), true_fp32():
        lo, ro = _oracle_lr_on_device(sdc, c, d, 2.0, 0.5)
    sl, sr = stats(l, lo), stats(r, ro)
    log_metric("row_flow_1080p", left_max=sl["max"], right_max=sr["max"], left_mean=sl["mean"], right_mean=sr["mean"])
    assert sl["mean"] < 1e-3 and sr["mean"] < 1e-3


@pytest.mark.parametrize("lk,rk,div,conv,sv,steps,border", [
    ("s2_left", "s2_right", 2.0, 0.5, "both", 2, False), ("s3b_left", "s3b_right", 4.0, 0.4, "both", 3, True),
    ("b_left", "b_right", 5.0, 0.5, "left", 1, True)])
def test_apply_divergence_nn_LR_steps_and_border_golden(lk, rk, div, conv, sv, steps, border):
    """steps > 1 (the depth is re-warped by each step's delta, iw3/backward_warp.py:205-226) and preserve_screen_border
    (:33-47) against the REAL reference model's fp32 CPU output (tests/golden/row_flow_steps.npz)."""
    from nunif_b200.iw3 import RowFlowV3, apply_divergence_nn_LR
    g = load_golden("row_flow_steps")
    m = RowFlowV3(synth.row_flow_v3_state_dict(0), DEV)
    c, d = t(g["c"], DEV), t(g["d"], DEV)
    l, r = apply_divergence_nn_LR(m, c, d, div, conv, steps=steps, synthetic_view=sv, preserve_screen_border=border)
    sl, sr = stats(l, t(g[lk])), stats(r, t(g[rk]))
    log_metric(f"row_flow_steps{steps}_border{int(border)}", left_max=sl["max"], right_max=sr["max"], left_mean=sl["mean"], right_mean=sr["mean"])
    # same bound as the single-step test: fp16 delta network against the reference's fp32 run, errors accumulate over the steps
    assert sl["mean"] < 1e-3 * steps and sr["mean"] < 1e-3 * steps and sl["max"] < 3e-2 * steps and sr["max"] < 3e-2 * steps, (sl, sr)
    if sv == "left":
        assert torch.equal(r, c)
