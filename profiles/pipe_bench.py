"""Where does the host-frame pipeline lose time?  Times (a) raw pinned H2D / D2H of one 1080p batch, (b) FrameBatchPipeline with a
trivial callback, (c) with the real depth + warp callback, (d) the per-call host overhead of pipe(frame)."""
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import synth  # noqa: E402
from nunif_b200.iw3 import stereo_sbs, DepthAnythingModel  # noqa: E402
from nunif_b200.nunif.video import FrameBatchPipeline  # noqa: E402

dev = torch.device("cuda:0")
B, H, W = 4, 1080, 1920
u8 = torch.randint(0, 255, (B, H, W, 3), dtype=torch.uint8).pin_memory()
d_in = torch.empty_like(u8, device=dev)
d_out = torch.empty((B, H, 2 * W, 3), dtype=torch.uint8, device=dev)
h_out = torch.empty((B, H, 2 * W, 3), dtype=torch.uint8).pin_memory()


def timed(fn, n=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print(f"H2D {u8.numel() / 1e6:.1f} MB batch: {timed(lambda: d_in.copy_(u8, non_blocking=True)):.3f} ms")
print(f"H2D per frame x4: {timed(lambda: [d_in[i].copy_(u8[i], non_blocking=True) for i in range(B)]):.3f} ms")
print(f"D2H {h_out.numel() / 1e6:.1f} MB batch: {timed(lambda: h_out.copy_(d_out, non_blocking=True)):.3f} ms")
dm = DepthAnythingModel().load_state_dict(synth.depth_anything_v2_state_dict(0), gpu=0)
c = torch.rand(B, 3, H, W, device=dev)


def real(xf):
    depth = dm.infer(xf, edge_dilation=[2, 1])
    return stereo_sbs(xf, depth, 2.0, 0.5, method="forward_fill", edge_dilation=0)


def trivial(xf):
    return torch.cat([xf, xf], dim=3)


with torch.inference_mode():
    print(f"real callback, device resident: {timed(lambda: real(c)):.3f} ms per batch")
for name, cb in (("trivial", trivial), ("real", real)):
    for depth in (3, 4):
        pipe = FrameBatchPipeline(cb, B, dev, depth=depth, copy_output=False)
        frames = [u8[i] for i in range(B)]
        for i in range(4 * B):
            pipe(frames[i % B])
        torch.cuda.synchronize()
        n = 40
        t0 = time.perf_counter()
        th = 0.0
        for i in range(n * B):
            t1 = time.perf_counter()
            pipe(frames[i % B])
            th += time.perf_counter() - t1
        pipe.finish()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"pipeline[{name}, depth {depth}]: {n * B / dt:.0f} fps, {dt / n * 1e3:.3f} ms per batch, host time inside pipe(): {th / n * 1e3:.3f} ms per batch")
