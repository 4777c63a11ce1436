"""Per-role clock64 timeline of CTA 0 of the fused qkv + window attention kernel (csrc/swin_fused_attn.cu FTL events)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib  # noqa: E402

DEV = "cuda:0"
L = _lib.lib()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 192
B, H = 16, 240
g = torch.Generator().manual_seed(2)
x = torch.randn(B, H, H, C, generator=g).half().to(DEV)
wqkv = (torch.randn(3 * C, C, generator=g) / C ** 0.5).half().to(DEV)
bqkv = torch.zeros(3 * C, device=DEV)
table = (0.5 * torch.randn(121, 6, generator=g)).to(DEV)
att = torch.empty(B, H, H, C, dtype=torch.float16, device=DEV)


def run():
    _lib.check(L.nb200_swin_attn_fused_f16(_lib.ptr(x), _lib.ptr(wqkv), _lib.ptr(bqkv), _lib.ptr(table), _lib.ptr(att), B, H, H, C, 3,
                                           _lib.stream_ptr()))


for _ in range(2):
    run()
torch.cuda.synchronize()
tl = torch.zeros(8 * 2048, dtype=torch.int64, device=DEV)
_lib.check(L.nb200_debug_timeline(_lib.ptr(tl)))
run()
torch.cuda.synchronize()
_lib.check(L.nb200_debug_timeline(None))
ev = tl.cpu().tolist()
NAMES = {1: "x wait-empty", 2: "x issue", 3: "w wait-empty", 4: "w issue", 10: "mma wait x", 11: "x ok", 12: "d buffer free", 13: "mma chunk issue",
         20: "epi wait d", 21: "d ok", 22: "qkv buffer free", 23: "epi done", 30: "attn wait qkv", 31: "qkv ok", 32: "attn math done", 33: "attn stored"}
TRACK = ["prodW", "prodX", "mma", "epi", "attn0", "attn2"]
rows = []
for tr in range(6):
    for i in range(2048):
        v = ev[tr * 2048 + i] & ((1 << 64) - 1)
        if v == 0:
            break
        rows.append((v & 0xffffffffff, tr, (v >> 56) & 0xff, (v >> 40) & 0xffff))
rows.sort()
starts = [t for (t, tr, tag, aux) in rows if tr == 2 and tag == 10]
print(f"C={C}: {len(starts)} tiles on CTA 0; cycles per tile (mma track):", [starts[i + 1] - starts[i] for i in range(min(len(starts) - 1, 12))])
if len(starts) > 8:
    t0, t1 = starts[5], starts[7]
    for (t, tr, tag, aux) in rows:
        if t0 <= t <= t1:
            print(f"{t - t0:8d}  {TRACK[tr]:6s} {NAMES.get(tag, tag)} [{aux}]")
