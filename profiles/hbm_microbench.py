"""HBM microbenchmarks used to interpret the rooflines: pure read, pure write, copy, and a 1:3 read:write mix
(the qkv GEMM's traffic shape).  Device-timed with CUDA events; buffers are far larger than the 126 MB L2."""
import json
import torch

dev = "cuda:0"
n = 1 << 30  # elements (2 GiB of fp16)
a = torch.empty(n, dtype=torch.float16, device=dev).normal_()
b = torch.empty(n, dtype=torch.float16, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3


res = {}
t = timeit(lambda: b.fill_(1.0)); res["write_only_GBps"] = n * 2 / t / 1e9
t = timeit(lambda: torch.cuda.memset if False else b.zero_()); res["memset_GBps"] = n * 2 / t / 1e9
t = timeit(lambda: a.sum()); res["read_only_GBps"] = n * 2 / t / 1e9
t = timeit(lambda: b.copy_(a)); res["copy_GBps"] = 2 * n * 2 / t / 1e9
# 1:3 read:write: write 3 outputs from 1 input
c = torch.empty((3, n // 4), dtype=torch.float16, device=dev)
src = a[: n // 4]
t = timeit(lambda: c.copy_(src.unsqueeze(0).expand(3, -1))); res["read1_write3_GBps"] = 4 * (n // 4) * 2 / t / 1e9
print(json.dumps(res))
