"""Mirror of waifu2x/utils.py:42-297 (class Waifu2x): model slots per method/noise level,
load_model (2x/1x derived from 4x), render, convert."""
from os import path
import torch
from ..nunif.models import load_model
from ..nunif.render import tiled_render
from ..nunif.alpha import AlphaBorderPadding
from ..nunif.tta import tta_split, tta_merge
import torch.nn.functional as F
from .. import _lib


class Waifu2x():
    def __init__(self, model_dir, gpus):
        self.scale_model = None
        self.scale4x_model = None
        self.noise_models = [None] * 4
        self.noise_scale_models = [None] * 4
        self.noise_scale4x_models = [None] * 4
        self.alpha_pad = AlphaBorderPadding()
        # nunif/device.py:12-32 create_device: gpus[0] < 0 means CPU in the reference; this engine is CUDA-only
        gpu = gpus[0] if isinstance(gpus, (list, tuple)) else gpus
        if isinstance(gpu, int) and gpu < 0:
            gpu = 0
        self.device = torch.device(f"cuda:{gpu}") if isinstance(gpu, int) else torch.device(gpu)
        if self.device.type != "cuda":
            raise RuntimeError("nunif_b200.waifu2x runs on CUDA (sm_100) only")
        self.gpus = gpus
        self.model_dir = model_dir
        self.is_half = False

    # API parity no-ops: the engine is already ahead-of-time compiled fp16
    def compile(self):
        return self

    def warmup(self, *args, **kwargs):
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("nunif_b200 models cannot be moved to the CPU")
        return self

    def half(self):
        self.is_half = True
        return self

    def float(self):
        self.is_half = False
        return self

    def load_model_by_name(self, filename):
        return load_model(path.join(self.model_dir, filename), device=self.device, weights_only=True)[0]

    def has_model_file(self, filename):
        return path.exists(path.join(self.model_dir, filename))

    def _load_model(self, method, noise_level):
        """waifu2x/utils.py:128-176."""
        if method == "scale4x":
            if self.scale4x_model is not None:
                return
            if self.has_model_file("scale4x.pth"):
                self.scale4x_model = self.load_model_by_name("scale4x.pth")
            else:
                raise FileNotFoundError(f"scale4x.pth not found in {self.model_dir}")
        elif method == "scale":
            if self.scale_model is not None:
                return
            if self.has_model_file("scale2x.pth"):
                self.scale_model = self.load_model_by_name("scale2x.pth")
            else:
                if self.scale4x_model is None:
                    self._load_model("scale4x", noise_level)
                self.scale_model = self.scale4x_model.to_2x()
        elif method == "noise_scale4x":
            if self.noise_scale4x_models[noise_level] is not None:
                return
            if self.has_model_file(f"noise{noise_level}_scale4x.pth"):
                self.noise_scale4x_models[noise_level] = self.load_model_by_name(f"noise{noise_level}_scale4x.pth")
            else:
                raise FileNotFoundError(f"noise{noise_level}_scale4x.pth not found in {self.model_dir}")
        elif method == "noise_scale":
            if self.noise_scale_models[noise_level] is not None:
                return
            if self.has_model_file(f"noise{noise_level}_scale2x.pth"):
                self.noise_scale_models[noise_level] = self.load_model_by_name(f"noise{noise_level}_scale2x.pth")
            else:
                if self.noise_scale4x_models[noise_level] is None:
                    self._load_model("noise_scale4x", noise_level)
                self.noise_scale_models[noise_level] = self.noise_scale4x_models[noise_level].to_2x()
        elif method == "noise":
            if self.noise_models[noise_level] is not None:
                return
            if self.has_model_file(f"noise{noise_level}.pth"):
                self.noise_models[noise_level] = self.load_model_by_name(f"noise{noise_level}.pth")
            else:
                if self.noise_scale4x_models[noise_level] is None:
                    self._load_model("noise_scale4x", noise_level)
                self.noise_models[noise_level] = self.noise_scale4x_models[noise_level].to_1x()
        else:
            raise ValueError(method)

    def load_model(self, method, noise_level):
        """waifu2x/utils.py:178-199."""
        assert (method in ("scale", "noise_scale", "noise", "scale4x", "noise_scale4x"))
        assert (method in {"scale", "scale4x"} or 0 <= noise_level and noise_level < 4)
        if method in {"scale", "scale4x", "noise"}:
            self._load_model(method, noise_level)
        elif method == "noise_scale4x":
            self._load_model(method, noise_level)
            try:
                self._load_model("scale4x", -1)
            except FileNotFoundError:
                pass
        elif method == "noise_scale":
            self._load_model(method, noise_level)
            try:
                self._load_model("scale", -1)
            except FileNotFoundError:
                pass

    def load_model_all(self, load_4x=True):
        """waifu2x/utils.py:201-216."""
        if load_4x:
            self._load_model("scale4x", -1)
            for noise_level in range(4):
                self._load_model("noise_scale4x", noise_level)
        self._load_model("scale", -1)
        for noise_level in range(4):
            self._load_model("noise_scale", noise_level)
        for noise_level in range(4):
            self._load_model("noise", noise_level)

    def _model(self, method, noise_level):
        return {"scale": lambda: self.scale_model, "scale4x": lambda: self.scale4x_model,
                "noise": lambda: self.noise_models[noise_level],
                "noise_scale": lambda: self.noise_scale_models[noise_level],
                "noise_scale4x": lambda: self.noise_scale4x_models[noise_level]}[method]()

    @torch.inference_mode()
    def render(self, x, method, noise_level, tile_size=None, batch_size=None, enable_amp=False):
        """waifu2x/utils.py:218-241."""
        assert (method in ("scale", "noise_scale", "noise", "scale4x", "noise_scale4x"))
        assert (method in {"scale", "scale4x"} or 0 <= noise_level and noise_level < 4)
        return tiled_render(x, self._model(method, noise_level), tile_size=tile_size, batch_size=batch_size,
                            enable_amp=enable_amp)

    def convert(self, x, alpha, method, noise_level, tile_size=None, batch_size=None,
                tta=False, enable_amp=False, output_device="cpu"):
        """waifu2x/utils.py:255-297."""
        assert (not torch.is_grad_enabled())
        assert (x.shape[0] == 3)
        assert (alpha is None or alpha.shape[0] == 1 and alpha.shape[1:] == x.shape[1:])
        assert (method in ("scale", "scale4x", "noise_scale", "noise_scale4x", "noise"))
        assert (method in {"scale", "scale4x"} or 0 <= noise_level and noise_level < 4)
        x = x.to(self.device)
        blank_alpha = True
        if alpha is not None:
            # check all 1 alpha channel (waifu2x/utils.py:266-268)
            blank_alpha = bool(torch.equal(alpha, torch.ones(alpha.shape, device=alpha.device, dtype=alpha.dtype)))
        if alpha is not None and not blank_alpha:
            alpha = alpha.to(self.device)
            x = self.alpha_pad(x, alpha, self._model(method, noise_level).i2i_offset)       # :269-271
        if tta:
            rgb = tta_merge([self.render(xx, method, noise_level, tile_size, batch_size, enable_amp)
                             for xx in tta_split(x)])                                         # :272-275
        else:
            rgb = self.render(x, method, noise_level, tile_size, batch_size, enable_amp)
        rgb = rgb.to(output_device)
        if alpha is not None and method in ("scale", "noise_scale", "scale4x", "noise_scale4x"):
            s = 4 if method in {"scale4x", "noise_scale4x"} else 2
            if not blank_alpha:
                model = self.scale4x_model if method in {"scale4x", "noise_scale4x"} else self.scale_model
                if model is not None:
                    # second render pass on the alpha plane (:282-287)
                    alpha = alpha.expand(3, alpha.shape[1], alpha.shape[2])
                    alpha = tiled_render(alpha, model, tile_size=tile_size, batch_size=batch_size,
                                         enable_amp=enable_amp).mean(0, keepdim=True)
                else:
                    alpha = F.interpolate(alpha.unsqueeze(0), scale_factor=s, mode="bilinear").squeeze(0)   # :288-291
            else:
                # all-ones alpha: the nearest upscale of ones is ones (:292-294)
                alpha = torch.ones((1, alpha.shape[1] * s, alpha.shape[2] * s), dtype=alpha.dtype, device=output_device)
            alpha = alpha.to(output_device)
        return rgb, alpha
