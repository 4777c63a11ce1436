"""Mirror of waifu2x/hub.py:10-175 (Waifu2xImageModel and the `waifu2x` hub factory)."""
from os import path
import os
import torch
from .utils import Waifu2x

MODEL_DIR = os.environ.get("NUNIF_B200_MODEL_DIR",
                           path.join(os.environ.get("NUNIF_HOME", path.expanduser("~/.nunif")), "waifu2x", "pretrained_models"))


def _types(model_dir):
    return {
        "art": path.join(model_dir, "swin_unet", "art"),
        "art_scan": path.join(model_dir, "swin_unet", "art_scan"),
        "photo": path.join(model_dir, "swin_unet", "photo"),
        "swin_unet/art": path.join(model_dir, "swin_unet", "art"),
        "swin_unet/art_scan": path.join(model_dir, "swin_unet", "art_scan"),
        "swin_unet/photo": path.join(model_dir, "swin_unet", "photo"),
        "cunet/art": path.join(model_dir, "cunet", "art"),
    }


MODEL_TYPES = _types(MODEL_DIR)
NO_4X_MODELS = {"cunet/art"}
METHODS = ["noise", "scale", "noise_scale", "scale2x", "noise_scale2x", "scale4x", "noise_scale4x"]


class Waifu2xImageModel():
    def __init__(self, model_type, method=None, noise_level=-1, device_ids=[-1], tile_size=None, batch_size=None,
                 keep_alpha=True, amp=True, model_dir=None):
        self.model_type = model_type
        self.tile_size = tile_size
        self.batch_size = batch_size
        self.keep_alpha = keep_alpha
        self.amp = amp
        types = _types(model_dir) if model_dir else MODEL_TYPES
        if model_type not in types:
            raise ValueError(f"model_type: choose from {list(types.keys())}")
        if method is not None and method not in METHODS:
            raise ValueError(f"method: choose from {METHODS}")
        if method is not None and method.startswith("noise") and noise_level not in {0, 1, 2, 3}:
            raise ValueError("noise_level: choose from [0, 1, 2, 3]")
        self.ctx = Waifu2x(types[model_type], device_ids)
        if method is not None:
            method = self.normalize_method(method, noise_level)
            self.ctx.load_model(method, noise_level)
            self.set_mode(method, noise_level)
        else:
            self.method = None
            self.noise_level = None
            self.ctx.load_model_all(load_4x=(model_type not in NO_4X_MODELS))

    def set_mode(self, method, noise_level=-1):
        method = self.normalize_method(method, noise_level)
        if self.model_type in NO_4X_MODELS and method in {"scale4x", "noise_scale4x"}:
            raise ValueError(f"method: {self.model_type} does not support {method}")
        if (method in {"noise", "noise_scale4x", "noise_scale", "noise_scale2x"} and noise_level not in {0, 1, 2, 3}):
            raise ValueError("noise_level: choose from (0, 1, 2, 3)")
        self.method = method
        self.noise_level = noise_level

    def compile(self):
        self.ctx.compile()
        return self

    def to(self, device):
        self.ctx = self.ctx.to(device)
        return self

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    def half(self):
        self.ctx.half()
        return self

    def float(self):
        self.ctx.float()
        return self

    @property
    def is_half(self):
        return self.ctx.is_half

    @property
    def device(self):
        return self.ctx.device

    def convert(self, input_filepath, output_filepath, tta=False, format="png", **kwargs):
        from PIL import Image
        new_im = self.infer_file(input_filepath, tta=tta, **kwargs)
        new_im.save(output_filepath, format=format)

    def infer_file(self, filepath, tta=False, output_type="pil", **kwargs):
        from PIL import Image
        return self.infer_pil(Image.open(filepath), tta=tta, output_type=output_type, **kwargs)

    def infer_pil(self, pil_image, tta=False, output_type="pil", **kwargs):
        import numpy as np
        has_alpha = pil_image.mode in ("RGBA", "LA") and self.keep_alpha
        arr = torch.from_numpy(np.asarray(pil_image.convert("RGBA" if has_alpha else "RGB"), dtype=np.uint8).copy())
        arr = arr.permute(2, 0, 1).float().div_(255.0)            # nunif/utils/pil_io.py:218-232 to_tensor
        rgb = arr[:3].to(self.device)
        alpha = arr[3:4].to(self.device) if has_alpha else None
        if self.is_half:
            rgb = rgb.half()
            alpha = alpha.half() if alpha is not None else None
        return self.infer_tensor(rgb, alpha, tta=tta, output_type=output_type, **kwargs)

    def infer_tensor(self, rgb, alpha=None, tta=False, output_type="pil", **kwargs):
        method = kwargs.get("method", self.method)
        noise_level = kwargs.get("noise_level", self.noise_level)
        if method is None:
            raise ValueError(("method is None. Call `model.set_mode(method, noise_level)`"
                              " or use method and noise_level kwargs"))
        method = self.normalize_method(method, noise_level if noise_level is not None else -1)
        with torch.inference_mode():
            rgb, alpha = self.ctx.convert(rgb, alpha, method, noise_level, tile_size=self.tile_size,
                                          batch_size=self.batch_size, tta=tta, enable_amp=self.amp)
        if output_type == "tensor":
            return (rgb, alpha)
        from PIL import Image
        x = rgb if alpha is None else torch.cat([rgb, alpha], dim=0)
        # pil_io.to_image: quantize256 = clamp(x*255 round) (nunif/utils/pil_io.py:235-253)
        x = torch.clamp(x.float() * 255.0, 0, 255).round().to(torch.uint8).permute(1, 2, 0).cpu().numpy()
        return Image.fromarray(x, mode="RGBA" if alpha is not None else "RGB")

    def infer(self, x, tta=False, output_type="pil", **kwargs):
        if isinstance(x, str):
            return self.infer_file(x, tta=tta, output_type=output_type, **kwargs)
        if torch.is_tensor(x):
            return self.infer_tensor(x, tta=tta, output_type=output_type, **kwargs)
        try:
            from PIL import Image
            if isinstance(x, Image.Image):
                return self.infer_pil(x, tta=tta, output_type=output_type, **kwargs)
        except ImportError:
            pass
        raise ValueError("Unsupported input format")

    def __call__(self, x, tta=False, output_type="pil", **kwargs):
        return self.infer(x, tta=tta, output_type=output_type, **kwargs)

    @staticmethod
    def normalize_method(method, noise_level):
        """waifu2x/hub.py:151-163."""
        if method is None:
            return None
        if method == "scale2x":
            method = "scale"
        if method == "noise_scale2x":
            method = "noise_scale"
        if method == "scale" and noise_level >= 0:
            method = "noise_scale"
        if method == "scale4x" and noise_level >= 0:
            method = "noise_scale4x"
        return method


def waifu2x(model_type="art", method=None, noise_level=-1, device_ids=[-1], tile_size=None, batch_size=None,
            keep_alpha=True, amp=True, **kwargs):
    """waifu2x/hub.py:166-175 without the network download (models must already be in MODEL_DIR)."""
    return Waifu2xImageModel(model_type=model_type, method=method, noise_level=noise_level, device_ids=device_ids,
                             tile_size=tile_size, batch_size=batch_size, keep_alpha=keep_alpha, amp=amp,
                             model_dir=kwargs.get("model_dir"))
