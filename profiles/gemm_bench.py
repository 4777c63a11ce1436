"""Micro-benchmark of the Swin Linear shapes of one 16-tile batch (T = 921600 tokens) through the C ABI.
Prints microseconds and the fraction of the traffic-mix HBM floor max((R+W)/copy, W/write_only, R/read_only)."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = "cuda:0"
T = 921600
COPY, WR, RD = 6.6e12, 7.4e12, 7.0e12   # streaming kernels, profiles/r1/hbm_mix.json (write-only 7.47, read+2 writes 6.86 TB/s)


def gemm(A, W, bias, out, act=0, res=None):
    M, K = A.shape
    N = W.shape[0]
    _lib.check(lib.nb200_conv_gemm_f16(_lib.ptr(A), 1, 1, M, K, K, 0, _lib.ptr(W), N, _lib.ptr(bias), act, _lib.ptr(out), out.shape[-1],
                                       0, 0, _lib.ptr(res), res.shape[-1] if res is not None else 0, 1, M, 0, 0, 0, _lib.stream_ptr()))


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
    return e0.elapsed_time(e1) / iters * 1e3


def floor_us(R, Wb):
    return max((R + Wb) / COPY, Wb / WR, R / RD) * 1e6


shapes = [("qkv", 192, 576, 0, False), ("proj", 192, 192, 0, True), ("fc1", 192, 384, 2, False), ("fc2", 384, 192, 0, True),
          ("qkv96", 96, 288, 0, False), ("fc1_96", 96, 192, 2, False), ("toimg", 192, 48, 0, False)]
variants = [("default", {}), ("bn256", {4: 256}), ("bn192", {4: 192}), ("bn128", {4: 128}), ("bn96", {4: 96}), ("bn64", {4: 64}),
            ("bn128_nq3", {4: 128, 0: 3})]


def check(A, W, b, out, act, res):
    """max |error| of the last launch against torch fp32 on a strided sample of rows (guards the tuning variants)"""
    idx = torch.arange(0, A.shape[0], 997, device=A.device)
    ref = A[idx].float() @ W.float().t() + b
    if act == 2:
        ref = torch.nn.functional.gelu(ref)
    if res is not None:
        ref = ref + res[idx].float()
    return float((out[idx].float() - ref).abs().max())

res_all = {}
for name, K, N, act, use_res in shapes:
    A = torch.randn(T, K, device=dev).half()
    W = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev)
    out = torch.empty(T, N, device=dev, dtype=torch.float16)
    res = torch.randn(T, N, device=dev).half() if use_res else None
    R = T * K * 2 + (T * N * 2 if use_res else 0)
    Wb = T * N * 2
    row = {}
    for vname, kv in variants:
        for k, v in kv.items():
            lib.nb200_tune_set(k, v)
        us = timeit(lambda: gemm(A, W, b, out, act, res))
        row[vname] = round(us, 1)
        err = check(A, W, b, out, act, res)
        row.setdefault("max_err", 0.0)
        row["max_err"] = max(row["max_err"], round(err, 5))
        out.zero_()
        lib.nb200_tune_set(0, 4); lib.nb200_tune_set(1, 8); lib.nb200_tune_set(2, 0); lib.nb200_tune_set(4, 0)
    row["floor_us"] = round(floor_us(R, Wb), 1)
    row["frac_default"] = round(row["floor_us"] / row["default"], 3)
    res_all[name] = row
    print(name, json.dumps(row), flush=True)
    del A, W, out, res
json.dump(res_all, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "gemm_bench.json"), "w"), indent=1)
