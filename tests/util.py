import os
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return {k: v for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}


def t(a, device="cpu"):
    return torch.from_numpy(np.asarray(a)).to(device)


def log_metric(name, **kv):
    """Append a line to gpurun_out/metrics.log (brought back from the GPU box)."""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "metrics.log"), "a") as fh:
            fh.write(name + " " + " ".join(f"{k}={v}" for k, v in kv.items()) + "\n")
    except OSError:
        pass


def stats(got, want):
    """max / mean / 99.9th percentile of |got - want| and the fraction above 1e-3.  Small tensors: float64 on the CPU; large ones
    (whole 4K x 4 frames, 4e8 samples) on the GPU in float32 with a float64 mean."""
    if got.numel() > (1 << 24) and torch.cuda.is_available():
        dev = got.device if got.is_cuda else (want.device if want.is_cuda else torch.device("cuda:0"))
        d = (got.to(dev, torch.float32) - want.to(dev, torch.float32)).abs().flatten()
        k = max(1, int(d.numel() * 0.999))
        return dict(max=float(d.max()), mean=float(d.sum(dtype=torch.float64) / d.numel()), p999=float(d.kthvalue(k).values),
                    frac_gt_1e3=float((d > 1e-3).sum(dtype=torch.float64) / d.numel()))
    d = (got.double().cpu() - want.double().cpu()).abs()
    return dict(max=float(d.max()), mean=float(d.mean()), p999=float(d.flatten().kthvalue(max(1, int(d.numel() * 0.999))).values),
                frac_gt_1e3=float((d > 1e-3).double().mean()))


import contextlib


@contextlib.contextmanager
def true_fp32():
    """The "fp32 reference" evaluated on the GPU must not use TF32 (cuDNN convolutions allow it by default: 10-bit mantissas, an
    error of the same size as the fp16 errors the parity tests measure)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
