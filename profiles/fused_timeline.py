"""Per-role clock64 timeline of CTA 0 of the fused MLP kernel (csrc/swin_fused_mlp.cu FTL events).
Usage (GPU box): python profiles/fused_timeline.py [C] > gpurun_out/fused_timeline_C.txt"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib  # noqa: E402

DEV = "cuda:0"
L = _lib.lib()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 192
T = 921600
g = torch.Generator().manual_seed(1)
x = torch.randn(T, C, generator=g).half().to(DEV)
att = torch.randn(T, C, generator=g).half().to(DEV)
wp = (torch.randn(C, C, generator=g) / C ** 0.5).half().to(DEV)
w1 = (torch.randn(2 * C, C, generator=g) / C ** 0.5).half().to(DEV)
w2 = (torch.randn(C, 2 * C, generator=g) / (2 * C) ** 0.5).half().to(DEV)
bp, b1, b2 = torch.zeros(C, device=DEV), torch.zeros(2 * C, device=DEV), torch.zeros(C, device=DEV)


def run():
    _lib.check(L.nb200_swin_mlp_fused_f16(_lib.ptr(x), _lib.ptr(att), T, C, _lib.ptr(wp), _lib.ptr(bp), _lib.ptr(w1), _lib.ptr(b1),
                                          _lib.ptr(w2), _lib.ptr(b2), _lib.stream_ptr()))


for _ in range(2):
    run()
torch.cuda.synchronize()
tl = torch.zeros(8 * 2048, dtype=torch.int64, device=DEV)
_lib.check(L.nb200_debug_timeline(_lib.ptr(tl)))
run()
torch.cuda.synchronize()
_lib.check(L.nb200_debug_timeline(None))
ev = tl.cpu().tolist()
NAMES = {1: "w wait-empty", 2: "w issue", 3: "act wait-empty", 4: "act issue", 10: "G0 wait att", 11: "G0 chunk", 12: "wait x1_ready",
         13: "x1_ready", 20: "G1 wait w", 21: "G1 issue", 30: "G2 wait h", 31: "G2 h ok", 32: "G2 issue", 40: "tile wait x", 41: "x ok",
         42: "d0 ok", 43: "E0 done", 50: "E1 wait d1", 51: "d1 ok", 52: "E1 math done / wait h_empty", 53: "h_empty ok", 60: "E2 wait d2",
         61: "d2 ok", 62: "store done"}
TRACK = ["prod0", "prod1", "prod2", "mma", "epi"]
rows = []
for tr in range(5):
    for i in range(2048):
        v = ev[tr * 2048 + i]
        if v == 0:
            break
        v &= (1 << 64) - 1
        rows.append((v & 0xffffffffff, tr, (v >> 56) & 0xff, (v >> 40) & 0xffff))
rows.sort()
# steady state: from the 6th "tile wait x" of the epilogue track for 2 tiles
starts = [t for (t, tr, tag, aux) in rows if tr == 4 and tag == 40]
print(f"C={C}: {len(starts)} tiles on CTA 0; cycles per tile (epilogue track):", [starts[i + 1] - starts[i] for i in range(min(len(starts) - 1, 12))])
if len(starts) > 8:
    t0, t1 = starts[5], starts[7]
    for (t, tr, tag, aux) in rows:
        if t0 <= t <= t1:
            print(f"{t - t0:8d}  {TRACK[tr]:6s} {NAMES.get(tag, tag)} [{aux}]")
