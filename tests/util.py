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
    d = (got.double().cpu() - want.double().cpu()).abs()
    return dict(max=float(d.max()), mean=float(d.mean()), p999=float(d.flatten().kthvalue(max(1, int(d.numel() * 0.999))).values),
                frac_gt_1e3=float((d > 1e-3).double().mean()))
