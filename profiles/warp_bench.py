"""Device-timed iw3 warp kernels at 1080p (depth 392x686) through the C ABI: back-to-back launches, CUDA events.
Algorithmic bytes per frame: 3 planes in + 6 planes out (fp32) + the depth map."""
import json
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib, synth  # noqa: E402
from nunif_b200.iw3 import apply_divergence_grid_sample, apply_divergence_forward_warp  # noqa: E402

lib = _lib.lib()
dev = "cuda:0"
H, W, h, w = 1080, 1920, 392, 686
COPY, WR, RD = 6.6e12, 7.4e12, 7.0e12   # streaming kernels, profiles/r1/hbm_mix.json (write-only 7.47, read+2 writes 6.86 TB/s)


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


res = {}
for B in (1, 4, 16):
    c = torch.stack([synth.synth_image(50 + i, 3, H, W, smooth=False) for i in range(min(B, 4))]).repeat((B + 3) // 4, 1, 1, 1)[:B].to(dev).contiguous()
    d = synth.synth_depth(60, B, h, w).to(dev)
    R = B * (3 * H * W * 4 + h * w * 4)
    Wb = B * 6 * H * W * 4
    floor = max((R + Wb) / COPY, Wb / WR, R / RD) * 1e6
    row = {"floor_us": round(floor, 1), "bytes": R + Wb}
    out = torch.empty((B, 3, H, 2 * W), device=dev)
    st = _lib.stream_ptr(dev)
    for name, tune in (("row_staged", 0), ("gather", 1)):
        lib.nb200_tune_set(3, tune)
        us = timeit(lambda: lib.nb200_backward_warp(_lib.ptr(c), _lib.ptr(d), B, H, W, h, w, 2.0, 0.5, 0, 1, _lib.ptr(out), None, st))
        row["backward_" + name + "_us"] = round(us, 1)
        row["backward_" + name + "_frac_of_floor"] = round(floor / us, 3)
        row["backward_" + name + "_frac_of_copy_peak"] = round((R + Wb) / (us * 1e-6) / 6558.1e9, 3)
    lib.nb200_tune_set(3, 0)
    ws = torch.empty(max(16, lib.nb200_forward_warp_workspace(B, H, W, h, w)), dtype=torch.uint8, device=dev)
    us = timeit(lambda: lib.nb200_forward_warp(_lib.ptr(c), _lib.ptr(d), B, H, W, h, w, 2.0, 0.5, 1, 0, 0, 1, _lib.ptr(out), None, None, None,
                                               _lib.ptr(ws), st))
    row["forward_fill_us"] = round(us, 1)
    row["forward_fill_frac_of_copy_peak"] = round((R + Wb) / (us * 1e-6) / 6558.1e9, 3)
    res[f"B{B}"] = row
    print(B, json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/warp_bench.json", "w"), indent=1)
