"""Render ONE synthetic frame through the engine (for ncu captures; never a bench value).
usage: python profiles/one_frame.py [4k|1080p] [warp|depth|zoe|upcunet]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import synth  # noqa: E402
from nunif_b200.nunif.models import create_model  # noqa: E402
from nunif_b200.nunif.render import tiled_render  # noqa: E402
from nunif_b200.iw3 import stereo_sbs  # noqa: E402

frame = sys.argv[1] if len(sys.argv) > 1 else "4k"
h, w = {"4k": (2160, 3840), "1080p": (1080, 1920)}[frame]
dev = "cuda:0"
if "warp" in sys.argv:
    c = synth.synth_image(1, 3, h, w, smooth=False).unsqueeze(0).to(dev)
    d = synth.synth_depth(2, 1, 392, 686).to(dev)
    for method in ("forward_fill", "backward"):
        for _ in range(3):
            y = stereo_sbs(c, d, 2.0, 0.5, method=method, edge_dilation=[2, 1])
    torch.cuda.synchronize()
elif "depth" in sys.argv:
    from nunif_b200.iw3 import DepthAnythingModel
    dam = DepthAnythingModel().load_state_dict(synth.depth_anything_v2_state_dict(0), gpu=0)
    c = torch.stack([synth.synth_image(50 + i, 3, h, w, smooth=False) for i in range(4)]).to(dev)
    with torch.inference_mode():
        for _ in range(2):
            y = dam.infer(c, edge_dilation=[2, 1])
    torch.cuda.synchronize()
elif "zoe" in sys.argv:
    from nunif_b200.iw3 import ZoeDepthModel
    zm = ZoeDepthModel("ZoeD_N").load_state_dict(synth.zoedepth_state_dict(0), gpu=0)
    c = torch.stack([synth.synth_image(50 + i, 3, h, w, smooth=False) for i in range(2)]).to(dev)
    with torch.inference_mode():
        for _ in range(2):
            d = zm.infer(c, edge_dilation=2)
            y = stereo_sbs(c, d, 2.0, 0.5, method="backward", mapper="div_6", edge_dilation=0, anaglyph="dubois")
    torch.cuda.synchronize()
elif "upcunet" in sys.argv:
    m = create_model("waifu2x.upcunet", synth.upcunet_state_dict(0), dev)
    x = synth.synth_image(1, 3, h, w, smooth=False).to(dev)
    with torch.no_grad():
        y = tiled_render(x, m, tile_size=256, batch_size=16)
    torch.cuda.synchronize()
else:
    m = create_model("waifu2x.swin_unet_4x", synth.swin_unet_state_dict(0, 4), dev)
    x = synth.synth_image(1, 3, h, w, smooth=False).to(dev)
    with torch.no_grad():
        y = tiled_render(x, m, tile_size=256, batch_size=16)
    torch.cuda.synchronize()
print("done", tuple(y.shape))
