"""GPU: the checkpoint-directory layer: Waifu2x.load_model* from `.pth` files in the
reference's format, 2x/1x derived from a 4x-only directory, Waifu2xImageModel.infer* and the error behaviour."""
import os
import pytest
import torch

from nunif_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _write(dirpath, filename, name, sd):
    os.makedirs(dirpath, exist_ok=True)
    # the dict nunif.models.save_model writes (nunif/models/utils.py:12-39)
    torch.save({"nunif_model": 1, "name": name, "kwargs": {}, "state_dict": sd, "updated_at": "0"}, os.path.join(dirpath, filename))


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory):
    root = tmp_path_factory.mktemp("pretrained_models")
    art = os.path.join(root, "swin_unet", "art")
    _write(art, "scale4x.pth", "waifu2x.swin_unet_4x", synth.swin_unet_state_dict(0, 4))
    _write(art, "noise1_scale4x.pth", "waifu2x.swin_unet_4x", synth.swin_unet_state_dict(1, 4))
    cu = os.path.join(root, "cunet", "art")
    _write(cu, "scale2x.pth", "waifu2x.upcunet", synth.upcunet_state_dict(0))
    _write(cu, "noise1.pth", "waifu2x.cunet", synth.cunet_state_dict(0))
    return str(root)


def test_waifu2x_directory_loading_and_derived_models(model_dir):
    from nunif_b200.waifu2x.utils import Waifu2x
    from nunif_b200.nunif.models import create_model
    from nunif_b200.nunif.render import tiled_render
    w = Waifu2x(os.path.join(model_dir, "swin_unet", "art"), [0])
    w.load_model("scale", -1)                       # no scale2x.pth: derived from scale4x.pth (waifu2x/utils.py:139-144)
    assert w.scale4x_model is not None and w.scale_model is not None and w.scale_model.i2i_scale == 2
    w.load_model("noise", 1)                        # noise1.pth missing: 4x + /4 (:164-170)
    assert w.noise_models[1].i2i_scale == 1 and w.noise_scale4x_models[1] is not None
    with pytest.raises(FileNotFoundError):
        w.load_model("noise_scale4x", 2)
    x = synth.synth_image(2, 3, 72, 100)
    with torch.inference_mode():
        rgb, alpha = w.convert(x, None, "scale", -1, tile_size=64, batch_size=4)
        want = tiled_render(x.to(DEV), create_model("waifu2x.swin_unet_4x", synth.swin_unet_state_dict(0, 4), DEV).to_2x(),
                            tile_size=64, batch_size=4)
    assert alpha is None and not rgb.is_cuda and torch.equal(rgb, want.cpu())


def test_hub_image_model(model_dir):
    from nunif_b200.waifu2x import waifu2x, Waifu2xImageModel
    m = waifu2x("art", method="scale4x", model_dir=model_dir, tile_size=64, batch_size=4)
    assert isinstance(m, Waifu2xImageModel) and m.device.type == "cuda"
    x = synth.synth_image(4, 3, 48, 64)
    rgb, alpha = m.infer(x, output_type="tensor")
    assert rgb.shape == (3, 192, 256) and alpha is None
    im = m.infer(x)                                  # PIL out
    assert im.size == (256, 192) and im.mode == "RGB"
    n = waifu2x("art", method="scale4x", noise_level=1, model_dir=model_dir, tile_size=64, batch_size=4)   # -> noise_scale4x (hub.py:151-163)
    assert n.method == "noise_scale4x" and n.infer(im, output_type="tensor")[0].shape == (3, 768, 1024)
    c = waifu2x("cunet/art", method="scale", model_dir=model_dir, tile_size=64, batch_size=4)
    assert c.infer(x, output_type="tensor")[0].shape == (3, 96, 128)
    with pytest.raises(ValueError):
        c.set_mode("scale4x")                        # hub.py:59-62
    with pytest.raises(ValueError):
        waifu2x("no_such_type", model_dir=model_dir)
    with pytest.raises(ValueError):
        waifu2x("art", method="noise", noise_level=7, model_dir=model_dir)
    with pytest.raises(ValueError):
        m.infer(3.14)
