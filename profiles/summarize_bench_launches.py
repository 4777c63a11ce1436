"""Summarise the ncu launch list of `bench.py --steps 1 --warmup 1` (profiles/r1/launches_bench_step.csv): per-kernel
share of the TIMED frame (the second frame of the list; frames end with the tile_gather_blend kernel) next to the
kernel-class shares bench.py measures live with CUDA events.  ncu times are serialised and cold-cache: shares, not
absolutes, are what must agree."""
import collections
import csv
import json
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "profiles/r1/launches_bench_step.csv"
bench = sys.argv[2] if len(sys.argv) > 2 else None
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
ends = [i for i, r in enumerate(rows) if "tile_gather_blend" in r[4]]
assert len(ends) >= 2, "need two complete frames in the list"
frame = rows[ends[0] + 1:ends[1] + 1]
agg = collections.OrderedDict()
tot = 0.0
for r in frame:
    name = r[4].split("(")[0].replace("void ", "").replace("nb200::", "")
    t = int(r[-1]) / 1e3
    tot += t
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
print(f"timed frame: {len(frame)} launches, sum of kernel durations {tot / 1e3:.2f} ms (serialised, cold caches)")
cls = collections.Counter()
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:48s} n={n:4d} {t / 1e3:8.3f} ms {100 * t / tot:5.1f}%")
    c = ("gemm" if "gemm_conv" in k else "fused_attn" if ("swin_attn_fused" in k or "swin_attn_tc" in k) else "fused_mlp" if "swin_mlp_fused" in k
         else "window_attention" if "window_attention" in k else "stem_conv" if "stem" in k
         else "to_image" if "to_image" in k else "tile_unfold" if "unfold" in k else "tile_blend" if "blend" in k else "other")
    cls[c] += t
print("class shares (ncu):   ", {k: round(v / tot, 3) for k, v in cls.most_common()})
if bench:
    d = json.loads(open(bench).read().strip().splitlines()[-1])
    kc = d["kernel_classes_ms"]
    s = sum(kc.values())
    print("class shares (bench): ", {k: round(v / s, 3) for k, v in sorted(kc.items(), key=lambda kv: -kv[1])})
