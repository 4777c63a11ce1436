"""Generate tests/golden/*.npz by running the REAL reference (nagadomi/nunif,
imported read-only from /root/reference) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):
    NUNIF_HOME=/tmp/nunif_home PYTHONDONTWRITEBYTECODE=1 \
    PYTHONPATH=/root/reference:/root/repo python oracle/gen_golden.py

The committed .npz files pin the oracle (tests/test_oracle_golden.py) and are
the fixtures the GPU parity tests compare against.  Inputs are regenerated at
test time from nunif_b200.synth with the seeds stored in each file, and the
inputs themselves are stored too so a synth drift is detected.
"""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NUNIF_HOME", "/tmp/nunif_home")
sys.path.insert(0, "/root/reference")

from nunif_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)
torch.manual_seed(0)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def gen_seam_config():
    from nunif.utils.seam_blending import SeamBlending
    rows = []
    cases = []
    for (scale, offset, blend) in [(1, 28, 0), (2, 36, 0), (1, 8, 4), (2, 16, 8), (4, 32, 16), (1, 8, 16)]:
        for tile in (64, 112, 160, 256, 400):
            for (h, w) in [(1, 1), (17, 33), (220, 220), (240, 240), (256, 256), (300, 420), (1080, 1920),
                           (2160, 3840), (4320, 7680), (236, 237), (471, 12)]:
                if tile - (2 * -(-offset // scale) + -(-blend // scale)) <= 0:
                    continue
                p = SeamBlending.create_config((h, w), scale, offset, tile, blend)
                cases.append([h, w, scale, offset, tile, blend])
                rows.append([p["y_h"], p["y_w"], p["h_blocks"], p["w_blocks"], *p["pad"],
                             p["y_buffer_h"], p["y_buffer_w"], p["input_tile_step"], p["output_tile_step"]])
    filters = {}
    for (scale, offset, tile, blend) in [(4, 32, 64, 16), (2, 16, 64, 8), (1, 8, 64, 4), (1, 8, 64, 16)]:
        filters[f"filter_{scale}_{offset}_{tile}_{blend}"] = SeamBlending.create_blend_filter(scale, offset, tile, blend, 3)
    save("seam_config", cases=np.array(cases, dtype=np.int64), configs=np.array(rows, dtype=np.int64), **filters)

    # find_valid_tile_size for both validators (model.py:51-62)
    import waifu2x.models  # noqa
    from nunif.models import create_model
    up = create_model("waifu2x.upcunet")
    sw = create_model("waifu2x.swin_unet_4x")
    q = np.arange(64, 600)
    save("tile_size", query=q,
         cunet=np.array([up.find_valid_tile_size(int(t)) for t in q]),
         swin=np.array([sw.find_valid_tile_size(int(t)) for t in q]))


def gen_models():
    import waifu2x.models  # noqa
    from nunif.models import create_model
    from nunif.utils.render import tiled_render
    # --- cunet family
    for name, sdf, n in [("waifu2x.upcunet", synth.upcunet_state_dict, "upcunet"),
                         ("waifu2x.cunet", synth.cunet_state_dict, "cunet")]:
        m = create_model(name).eval()
        m.load_state_dict(sdf(0), strict=True)
        x = synth.synth_image(11, 3, 104, 104).unsqueeze(0)
        z = m(x)
        img = synth.synth_image(12, 3, 150, 170)
        y = tiled_render(img, m, tile_size=104, batch_size=3)
        save(n, x=x, z=z, img=img, render=y, tile_size=104, batch_size=3)
    # --- swin family
    m4 = create_model("waifu2x.swin_unet_4x").eval()
    m4.load_state_dict(synth.swin_unet_state_dict(0, 4), strict=True)
    x = synth.synth_image(21, 3, 64, 64).unsqueeze(0).repeat(2, 1, 1, 1)
    x[1] = synth.synth_image(22, 3, 64, 64)
    z4 = m4(x)
    m2 = m4.to_2x().eval()  # waifu2x/utils.py:99-100 (_setup -> .eval())
    z2 = m2(x)
    m1 = m4.to_1x().eval()
    z1 = m1(x)
    img = synth.synth_image(23, 3, 70, 100)
    y4 = tiled_render(img, m4, tile_size=64, batch_size=4)
    y2 = tiled_render(img, m2, tile_size=64, batch_size=4)
    save("swin_unet_4x", x=x, z4=z4, z2=z2, z1=z1, img=img, render4=y4, render2=y2, tile_size=64, batch_size=4)
    for sf, name in [(2, "waifu2x.swin_unet_2x"), (1, "waifu2x.swin_unet_1x")]:
        m = create_model(name).eval()
        m.load_state_dict(synth.swin_unet_state_dict(0, sf), strict=True)
        save(f"swin_unet_{sf}x", x=x[:1], z=m(x[:1]))


def gen_iw3():
    from iw3.backward_warp import apply_divergence_grid_sample
    from iw3.forward_warp import apply_divergence_forward_warp
    from iw3.dilation import dilate_edge
    from iw3.depth_scaler import minmax_normalize
    from iw3.mapper import get_mapper
    from iw3.anaglyph import apply_anaglyph_redcyan

    B, H, W, h, w = 2, 72, 128, 28, 49
    c = torch.stack([synth.synth_image(31 + i, 3, H, W) for i in range(B)])
    d_lo = synth.synth_depth(41, B, h, w)
    d_hi = synth.synth_depth(42, B, H, W)

    out = {}
    for sv in ("both", "left", "right"):
        l, r = apply_divergence_grid_sample(c, d_lo, 2.0, 0.5, sv)
        out[f"bw_{sv}_l"], out[f"bw_{sv}_r"] = l, r
    l, r = apply_divergence_grid_sample(c, d_hi, 5.0, 0.3, "both")
    out["bw_hi_l"], out["bw_hi_r"] = l, r
    save("backward_warp", c=c, d_lo=d_lo, d_hi=d_hi, **out)

    out = {}
    for tag, depth, div, conv, wb in [("hi", d_hi, 4.0, 0.5, False), ("hi_wb", d_hi, 10.0, 0.3, True),
                                      ("lo", d_lo, 4.0, 0.5, False)]:
        for method in ("forward_fill", "forward"):
            l, r, lm, rm = apply_divergence_forward_warp(c.clone(), depth.clone(), div, conv, method=method,
                                                         synthetic_view="both", return_mask=True, width_base=wb)
            out[f"fw_{tag}_{method}_l"], out[f"fw_{tag}_{method}_r"] = l, r
            out[f"fw_{tag}_{method}_lm"], out[f"fw_{tag}_{method}_rm"] = lm, rm
    for sv in ("left", "right"):
        l, r = apply_divergence_forward_warp(c.clone(), d_hi.clone(), 2.0, 0.5, method="forward_fill",
                                             synthetic_view=sv, width_base=False)
        out[f"fw_{sv}_l"], out[f"fw_{sv}_r"] = l, r
    # long holes: > 100 px (iteration cap, forward_warp.py:18,45)
    Wl = 1280
    cl = synth.synth_image(35, 3, 12, Wl).unsqueeze(0)
    dl = synth.synth_depth(45, 1, 12, Wl)
    l, r = apply_divergence_forward_warp(cl.clone(), dl.clone(), 60.0, 0.0, method="forward_fill",
                                         synthetic_view="both", width_base=True)
    out["fw_long_l"], out["fw_long_r"] = l, r
    save("forward_warp", c=c, d_lo=d_lo, d_hi=d_hi, cl=cl, dl=dl, **out)

    dd = synth.synth_depth(51, 2, 56, 98) * 7.0 + 0.25
    out = {"x": dd}
    for n in ([2, 1], [1, 2], 2, [0, 0], [3, 0]):
        key = "dil_" + "_".join(str(v) for v in (n if isinstance(n, list) else [n]))
        out[key] = dilate_edge(dd.clone(), n)
    mn, mx = dd[0].amin(), dd[0].amax()
    out["minmax0"] = minmax_normalize(dd[0].clone(), mn, mx)
    out["div_6"] = get_mapper("div_6")(out["minmax0"])
    out["div_1"] = get_mapper("div_1")(out["minmax0"])
    save("dilation", **out)

    le, re = synth.synth_image(61, 3, 40, 64), synth.synth_image(62, 3, 40, 64)
    out = {"l": le, "r": re}
    for t in ("dubois", "dubois2", "color", "gray", "half-color", "wimmer", "wimmer2"):
        out[t.replace("-", "_")] = apply_anaglyph_redcyan(le.clone(), re.clone(), t)
    save("anaglyph", **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["seam", "models", "iw3"]
    if "seam" in which:
        gen_seam_config()
    if "models" in which:
        gen_models()
    if "iw3" in which:
        gen_iw3()
