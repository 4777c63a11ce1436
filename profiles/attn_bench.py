"""Device-timed window attention (one 16-tile batch: 16 x 240 x 240 tokens) through the C ABI, with the shared-memory
carveout sweep.  Floor = traffic-mix HBM floor of q,k,v in + out (profiles/r1/hbm_microbench.json)."""
import json
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = "cuda:0"
COPY, WR, RD = 6.6e12, 7.4e12, 7.0e12   # streaming kernels, profiles/r1/hbm_mix.json (write-only 7.47, read+2 writes 6.86 TB/s)
B, H = 16, 240
T = B * H * H
res = {}
for C in (192, 96):
    qkv = torch.randn(3, T, C, device=dev).half()
    table = torch.randn(121, 6, device=dev) * 0.5
    out = torch.empty(T, C, device=dev, dtype=torch.float16)
    R, Wb = 3 * T * C * 2, T * C * 2
    floor = max((R + Wb) / COPY, Wb / WR, R / RD) * 1e6
    row = {"floor_us": round(floor, 1)}
    for shift in (0, 3):
        for carve in (0, 100, 86, 72, 58, 44):
            lib.nb200_tune_set(6, carve)

            def run():
                _lib.check(lib.nb200_window_attention_f16(_lib.ptr(qkv), _lib.ptr(table), _lib.ptr(out), B, H, H, C, 6, shift, _lib.stream_ptr()))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            row[f"shift{shift}_carve{carve}_us"] = round(e0.elapsed_time(e1) / 10 * 1e3, 1)
    lib.nb200_tune_set(6, 0)
    row["frac_of_floor_default"] = round(floor / row["shift3_carve0_us"], 3)
    res[f"C{C}"] = row
    print(C, json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/attn_bench.json", "w"), indent=1)
