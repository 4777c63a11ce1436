"""Bisect why FrameBatchPipeline is ~10x slower inside bench.py than standalone (profiles/pipe_bench.py): each variant runs in
its own process and adds one ingredient of bench.py's iw3 path before timing the trivial-callback pipeline."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = ["base", "dist_import", "set_device", "clock_sampler", "synth_frames", "device_loop_first", "check_device", "all"]


def child(v):
    sys.path.insert(0, ROOT)
    import torch
    if v in ("dist_import", "all"):
        import torch.distributed as dist  # noqa: F401
    from nunif_b200 import synth, _lib
    from nunif_b200.nunif.video import FrameBatchPipeline
    if v in ("set_device", "all"):
        torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    if v in ("check_device", "all"):
        _lib.check(_lib.lib().nb200_check_device(0))
    B, H, W = 4, 1080, 1920
    if v in ("clock_sampler", "all"):
        import bench
        with bench.ClockSampler(0):
            x = torch.rand(1 << 24, device=dev)
            for _ in range(50):
                x = x * 1.0001
            torch.cuda.synchronize()
            time.sleep(0.3)
    if v in ("device_loop_first", "all"):
        from nunif_b200.iw3 import stereo_sbs, DepthAnythingModel
        dm = DepthAnythingModel().load_state_dict(synth.depth_anything_v2_state_dict(0), gpu=0)
        c = torch.rand(B, 3, H, W, device=dev)
        with torch.inference_mode():
            for _ in range(10):
                y = stereo_sbs(c, dm.infer(c, edge_dilation=[2, 1]), 2.0, 0.5, method="forward_fill", edge_dilation=0)
        torch.cuda.synchronize()
    if v in ("synth_frames", "all"):
        c = torch.stack([synth.synth_image(50 + i, 3, H, W, smooth=False) for i in range(B)]).to(dev)
        u8 = (c.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).cpu().pin_memory()
    else:
        u8 = torch.randint(0, 255, (B, H, W, 3), dtype=torch.uint8).pin_memory()
    frames = [u8[i] for i in range(B)]
    pipe = FrameBatchPipeline(lambda xf: torch.cat([xf, xf], dim=3), B, dev, depth=3, copy_output=False)
    for i in range(4 * B):
        pipe(frames[i % B])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(40 * B):
        pipe(frames[i % B])
    pipe.finish()
    torch.cuda.synchronize()
    print(f"{v:20s} {40 * B / (time.perf_counter() - t0):8.0f} fps   contiguous={u8.is_contiguous()} stride={u8.stride()}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for v in VARIANTS:
            subprocess.run([sys.executable, os.path.abspath(__file__), v], timeout=120)
