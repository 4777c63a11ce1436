"""Per-launch roofline table of ONE 16-tile batch of the swin_unet 4x forward (library event timing):
time, algorithmic HBM bytes / 6.56 TB/s floor, fraction, and FLOP/B for the GEMM launches."""
import ctypes
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import synth, _lib  # noqa: E402
from nunif_b200.nunif.models import create_model  # noqa: E402

lib = _lib.lib()
dev = "cuda:0"
for k, v in (kv.split("=") for kv in os.environ.get("NB200_TUNE", "").split(",") if kv):   # e.g. NB200_TUNE=11=1,10=0
    lib.nb200_tune_set(int(k), int(v))
which = sys.argv[1] if len(sys.argv) > 1 else "swin4x"
if which == "upcunet":
    m = create_model("waifu2x.upcunet", synth.upcunet_state_dict(0), dev)
else:
    m = create_model("waifu2x.swin_unet_4x", synth.swin_unet_state_dict(0, 4), dev)
x = torch.rand(16, 3, 256, 256, device=dev)
for _ in range(3):
    z = m(x)
torch.cuda.synchronize()
lib.nb200_profile_enable(1)
z = m(x)
buf = ctypes.create_string_buffer(1 << 20)
_lib.check(lib.nb200_profile_dump(buf, 1 << 20))
lib.nb200_profile_enable(0)
rows = [r.split(",") for r in buf.value.decode().strip().splitlines()]
tot = sum(float(r[1]) for r in rows)
print(f"{len(rows)} launches, {tot:.3f} ms per 16-tile batch")
print(f"{'#':>3} {'class':18s} {'us':>8s} {'floor_us':>9s} {'frac':>6s} {'GB':>7s} {'FLOP/B':>7s} {'lost_us':>8s}")
lost_total = 0.0
for i, r in enumerate(rows):
    cat, ms, work, rb, wb = r[0], float(r[1]), float(r[2]), float(r[3]), float(r[4])
    by = rb + wb
    floor = by / 6.558e12 * 1e6
    us = ms * 1e3
    inten = work / by if by > 0 and cat == "gemm" else 0.0
    lost = us - floor if by > 0 else 0.0
    lost_total += max(lost, 0.0)
    print(f"{i:3d} {cat:18s} {us:8.1f} {floor:9.1f} {floor / us if us > 0 and by > 0 else 0:6.2f} {by / 1e9:7.3f} {inten:7.1f} {lost:8.1f}")
print(f"sum of (time - HBM floor) over launches with known bytes: {lost_total:.1f} us of {tot * 1e3:.1f}")
