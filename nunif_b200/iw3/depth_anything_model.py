"""Mirror of iw3/depth_anything_model.py (DepthAnythingModel / batch_infer, lines 113-253) for the
Depth-Anything-V2 ViT-S network on the B200 engine.

The reference obtains the network from torch.hub ("nagadomi/Depth-Anything_iw3:main", DepthAnything(encoder="v2_vits"),
depth_anything_model.py:223-230); here the same checkpoint (upstream key names ``pretrained.*`` / ``depth_head.*``) is
packed into the native container (csrc/depth_model.inl) and run as tcgen05 GEMMs + the kernels in
csrc/depth_kernels.cu.  ``infer`` keeps the reference's signature and output convention: depth B,1,h,w (or 1,h,w)
float32 on ``x.device``, larger = nearer.
"""
import ctypes
import torch
from .. import _lib
from .depth_anything_preprocess import batch_preprocess
from .dilation import dilate_edge, edge_dilation_is_enabled

# NB200_MODEL_DEPTH_ANYTHING_V2_{S,B,L}; model types as in iw3/depth_anything_model.py NAME_MAP
KINDS = {"vits": 6, "vitb": 8, "vitl": 9}
ENCODER_OF = {"Any_V2_S": "vits", "Any_V2_B": "vitb", "Any_V2_L": "vitl"}


class DepthAnythingNet:
    """The packed network: ``net(x)`` == ``DepthAnythingV2.forward`` (x: B,3,H,W normalised, H,W % 14 == 0 -> B,H,W)."""

    def __init__(self, state_dict, device="cuda:0", encoder="vits"):
        if encoder not in KINDS:
            raise ValueError(f"encoder: choose from {list(KINDS)}")
        self.encoder = encoder
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("nunif_b200 models live on a CUDA (sm_100) device; there is no CPU path")
        items = [(k, v.detach().to("cpu", torch.float32).contiguous()) for k, v in state_dict.items()]
        n = len(items)
        names = (ctypes.c_char_p * n)(*[k.encode() for k, _ in items])
        datas = (ctypes.c_void_p * n)(*[v.data_ptr() for _, v in items])
        numels = (ctypes.c_int64 * n)(*[v.numel() for _, v in items])
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nb200_model_create(KINDS[encoder], n, names, datas, numels, 0, ctypes.byref(h)))
        self._h = h
        self.metric_depth = False
        self.prep_lower_bound = 392

    def __del__(self):
        try:
            if self._h:
                _lib.lib().nb200_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def __call__(self, x):
        _lib.require_cuda(x, "x")
        assert x.ndim == 4 and x.shape[1] == 3
        B, _, H, W = x.shape
        xf = x.float().contiguous()
        out = torch.empty((B, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().nb200_depth_anything_forward(self._h, _lib.ptr(xf), B, H, W, _lib.ptr(out), _lib.stream_ptr(x.device)))
        return out


@torch.inference_mode()
def batch_infer(model, im, flip_aug=True, low_vram=False, enable_amp=False, output_device="cpu", device=None,
                edge_dilation=2, depth_aa=None, limit_resolution=False, **kwargs):
    """depth_anything_model.py:122-182.  ``enable_amp`` is accepted for signature parity: the engine always runs the
    reference's CUDA numerics (fp16 autocast).  ``depth_aa`` (a learned post-filter, iw3/models/depth_aa.py) is not part
    of the B200 path."""
    device = device if device is not None else model.device
    assert torch.is_tensor(im) and im.ndim in (3, 4)
    batch = im.ndim == 4
    x = (im if batch else im.unsqueeze(0)).to(device)
    if depth_aa is not None:
        raise NotImplementedError("depth_aa is not implemented by the B200 engine")
    x = batch_preprocess(x, model.prep_lower_bound, limit_resolution=limit_resolution)
    if flip_aug:
        x = torch.cat([x, torch.flip(x, dims=[3])], dim=0)           # :140-142 (low_vram only changes the batching)
    out = torch.nan_to_num(model(x).unsqueeze(1))                    # _forward :113-119
    if edge_dilation_is_enabled(edge_dilation):
        out = dilate_edge(out, edge_dilation) if not model.metric_depth else -dilate_edge(-out, edge_dilation)
    if model.metric_depth:
        out = -out
    if flip_aug:
        n = out.shape[0] // 2
        z = (out[:n] + torch.flip(out[n:], dims=[3])) * 0.5          # :163-171
    else:
        z = out
    if not batch:
        z = z.squeeze(0)
    return z.to(output_device)


class DepthAnythingModel:
    """BaseDepthModel-shaped wrapper (iw3/base_depth_model.py) around a DepthAnythingNet."""

    def __init__(self, model_type="Any_V2_S"):
        if model_type not in ENCODER_OF:
            raise ValueError(f"the B200 engine implements {list(ENCODER_OF)} (Depth-Anything-V2 relative-depth models)")
        self.model_type = model_type
        self.model = None
        self.device = None
        self.limit_resolution = False

    def load_state_dict(self, state_dict, gpu=0, resolution=None, limit_resolution=False):
        """The reference downloads the checkpoint through torch.hub (depth_anything_model.py:223-230); here the caller
        passes the same state_dict (e.g. torch.load of depth_anything_v2_vits.pth)."""
        self.device = torch.device(f"cuda:{gpu}") if isinstance(gpu, int) else torch.device(gpu)
        self.model = DepthAnythingNet(state_dict, self.device, encoder=ENCODER_OF[self.model_type])
        lb = resolution or 392                                        # :232-235
        if lb % 14 != 0:
            lb += 14 - lb % 14
        self.model.prep_lower_bound = lb
        self.limit_resolution = limit_resolution
        return self

    def is_metric(self):
        return False

    @classmethod
    def get_name(cls):
        return "DepthAnything"

    def infer(self, x, tta=False, low_vram=False, enable_amp=True, edge_dilation=0, depth_aa=False, **kwargs):
        """depth_anything_model.py:241-253."""
        _lib.require_cuda(x, "x")
        if depth_aa:
            raise NotImplementedError("depth_aa is not implemented by the B200 engine")
        return batch_infer(self.model, x, flip_aug=tta, low_vram=low_vram, enable_amp=enable_amp, output_device=x.device,
                           device=x.device, edge_dilation=edge_dilation, depth_aa=None,
                           limit_resolution=self.limit_resolution)
