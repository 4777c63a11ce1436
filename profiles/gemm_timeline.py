"""Dump the per-role clock64 timeline of CTA 0 for one persistent GEMM launch (debug instrumentation)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = "cuda:0"
T = 921600
shape = sys.argv[1] if len(sys.argv) > 1 else "qkv"
K, N, act, use_res = {"qkv": (192, 576, 0, False), "proj": (192, 192, 0, True), "fc1": (192, 384, 2, False),
                      "toimg": (192, 48, 0, False), "fc2": (384, 192, 0, True), "fc1_96": (96, 192, 2, False),
                      "qkv96": (96, 288, 0, False), "proj96": (96, 96, 0, True), "fc2_96": (192, 96, 0, True)}[shape]
if len(sys.argv) > 2:
    lib.nb200_tune_set(4, int(sys.argv[2]))   # forced BLOCK_N
A = torch.randn(T, K, device=dev).half()
W = (torch.randn(N, K, device=dev) / K ** 0.5).half()
b = torch.randn(N, device=dev)
out = torch.empty(T, N, device=dev, dtype=torch.float16)
res = torch.randn(T, N, device=dev).half() if use_res else None


def run():
    _lib.check(lib.nb200_conv_gemm_f16(_lib.ptr(A), 1, 1, T, K, K, 0, _lib.ptr(W), N, _lib.ptr(b), act, _lib.ptr(out), N, 0, 0,
                                       _lib.ptr(res), N if use_res else 0, 1, T, 0, 0, 0, _lib.stream_ptr()))


for _ in range(3):
    run()
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
lib.nb200_debug_timeline(_lib.ptr(buf))
run()
torch.cuda.synchronize()
lib.nb200_debug_timeline(None)
v = buf.cpu().numpy().astype("uint64")
t0 = min(int(x) & ((1 << 56) - 1) for x in v if x)
names = {1: "prod:slot_free", 2: "mma:acc_free", 3: "mma:a_full", 4: "epi:acc_full", 5: "epi:chunk_begin", 6: "epi:math_done",
         7: "epi:fence_done", 8: "epi:after_bar", 9: "epi:leader_done"}
for base, role in ((0, "PROD"), (1024, "MMA"), (2048, "QUAD0"), (3072, "QUAD1")):
    ev = [(int(x) >> 56, (int(x) & ((1 << 56) - 1)) - t0) for x in v[base:base + 1024] if x]
    print(f"== {role}: {len(ev)} events (first 24)")
    # steady-state mean gap BEFORE each event type (second half of the record)
    half = ev[len(ev) // 2:]
    gaps = {}
    for (e0, c0), (e1, c1) in zip(half[:-1], half[1:]):
        gaps.setdefault(names.get(e1, e1), []).append(c1 - c0)
    print("   steady mean clk before event:", {k: round(sum(g) / len(g)) for k, g in gaps.items()},
          "span/event-cycle:", round((half[-1][1] - half[0][1]) / max(1, sum(1 for e, _ in half if e == half[0][0]))))
    prev = 0
    for e, c in ev[:24]:
        print(f"   {names.get(e, e):16s} t={c:8d}  (+{c - prev})")
        prev = c
    if len(ev) > 200:
        # steady-state period: clock delta between the same event 30 occurrences apart
        same = [c for e, c in ev if e == ev[0][0]]
        print(f"   steady period per occurrence of first event: {(same[-1] - same[len(same) // 2]) / (len(same) - 1 - len(same) // 2):.1f} clk")
