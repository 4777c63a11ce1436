"""Stage-by-stage comparison of the ZoeDepth engine with the oracle (one GPU call diagnoses every stage).
usage: python profiles/debug_zoe.py [mini|full] [H W] [B]
Taps (nb200_debug_tap): 0-3 hooked hidden states, 4-7 path_4..path_1, 8 bottleneck, 9 out_conv activation, 10 relative
depth, 11-14 bin centres after each attractor layer; then the final metric depth."""
import ctypes
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import synth, _lib  # noqa: E402
from nunif_b200.iw3 import ZoeDepthNet  # noqa: E402
from oracle import zoedepth as oz  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "mini"
H = int(sys.argv[2]) if len(sys.argv) > 3 else (64 if mode == "mini" else 384)
W = int(sys.argv[3]) if len(sys.argv) > 3 else (96 if mode == "mini" else 512)
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
cfg_s, cfg_o = (synth.ZOED_MINI, oz.ZOED_MINI) if mode == "mini" else (synth.ZOED_N, oz.ZOED_N)
dev = "cuda:0"
sd = synth.zoedepth_state_dict(1, cfg_s)
x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(H + W))
t0 = time.time()
net = ZoeDepthNet(sd, dev)
print(f"pack + upload: {time.time() - t0:.1f} s")
sdc = {k: v.to(dev) for k, v in sd.items()}
torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False   # a true fp32 reference
with torch.no_grad():
    ref = oz.zoedepth_forward(sdc, x.to(dev), cfg_o, return_all=True)
    with torch.autocast("cuda", dtype=torch.float16):
        amp = oz.zoedepth_forward(sdc, x.to(dev), cfg_o, return_all=True)


def nhwc(t):   # oracle NCHW -> engine NHWC
    return t.permute(0, 2, 3, 1).contiguous()


dim, F = cfg_o["dim"], cfg_o["feat"]
ph, pw = H // 16, W // 16
stages = []
for i in range(4):
    stages.append((i, f"hook{i}", ref["feats"][i], amp["feats"][i], torch.float16))
for i in range(4):
    stages.append((4 + i, f"path{4 - i}", nhwc(ref["blocks"][i]), nhwc(amp["blocks"][i]), torch.float16))
stages.append((8, "bottleneck", nhwc(ref["bottleneck"]), nhwc(amp["bottleneck"]), torch.float16))
stages.append((9, "out_conv act", nhwc(ref["act"]), nhwc(amp["act"]), torch.float16))
stages.append((10, "relative depth", ref["rel"], amp["rel"], torch.float32))
for i in range(4):
    stages.append((11 + i, f"bins level {i}", nhwc(ref["bins"][i]), nhwc(amp["bins"][i]), torch.float32))
buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.lib()


def report(name, got, r32, ramp):
    got, r32, ramp = got.float().cpu(), r32.float().cpu(), ramp.float().cpu()
    sc = float(r32.abs().max())
    e = (got - r32).abs()
    ea = (ramp - r32).abs()
    print(f"{name:16s} scale {sc:9.4f}  ours-fp32 max {float(e.max()):.3e} mean {float(e.mean()):.3e} | amp-fp32 max {float(ea.max()):.3e} "
          f"mean {float(ea.mean()):.3e} | finite {bool(torch.isfinite(got).all())}", flush=True)


xd = x.to(dev)
for tid, name, r32, ramp, dt in stages:
    _lib.check(lib.nb200_debug_tap(tid, ctypes.c_void_p(buf.data_ptr()), buf.numel()))
    out = net(xd)
    torch.cuda.synchronize()
    n = r32.numel()
    got = buf[: n * (2 if dt == torch.float16 else 4)].view(dt)[:n].view(r32.shape).clone()
    report(name, got, r32, ramp)
_lib.check(lib.nb200_debug_tap(-1, None, 0))
out = net(xd)
torch.cuda.synchronize()
report("metric depth", out, ref["metric_depth"], amp["metric_depth"])
e = (out.float().cpu() - ref["metric_depth"].float().cpu()).abs()
flat = e.flatten().topk(5)
for v, idx in zip(flat.values.tolist(), flat.indices.tolist()):
    b, r = divmod(idx, H * W)
    yy, xx = divmod(r, W)
    print(f"  worst pixel b={b} y={yy} x={xx}: ours {float(out[b, 0, yy, xx]):.5f} fp32 {float(ref['metric_depth'][b, 0, yy, xx]):.5f} amp "
          f"{float(amp['metric_depth'][b, 0, yy, xx]):.5f} | p fp32 {float(ref['p'][b, yy, xx]):.6f} amp {float(amp['p'][b, yy, xx]):.6f} | "
          f"T fp32 {float(ref['temperature'][b, 0, yy, xx]):.5f} amp {float(amp['temperature'][b, 0, yy, xx]):.5f}")
ea = (amp["metric_depth"].float().cpu() - ref["metric_depth"].float().cpu()).abs()
flat = ea.flatten().topk(3)
for v, idx in zip(flat.values.tolist(), flat.indices.tolist()):
    b, r = divmod(idx, H * W)
    yy, xx = divmod(r, W)
    print(f"  worst AMP pixel b={b} y={yy} x={xx}: err {v:.5f} T fp32 {float(ref['temperature'][b, 0, yy, xx]):.5f}")
# timing
for _ in range(2):
    net(xd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    net(xd)
e1.record()
torch.cuda.synchronize()
print(f"forward B={B} {H}x{W}: {e0.elapsed_time(e1) / 5:.3f} ms")
