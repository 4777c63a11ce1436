"""swin_attn_tc.cu (tcgen05 QK^T / PV) against swin_fused_attn.cu (mma.sync attention warps): max difference on a shape sweep,
then the timing of both at the bench shapes.  Usage (GPU box): python profiles/attn_tc_check.py [out.json] [--timeline C]"""
import json
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib  # noqa: E402

DEV = "cuda:0"
L = _lib.lib()


def make(B, H, W, C, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, H, W, C, generator=g).half().to(DEV)
    wqkv = (torch.randn(3 * C, C, generator=g) / C ** 0.5).half().to(DEV)
    bqkv = (0.1 * torch.randn(3 * C, generator=g)).to(DEV)
    table = (0.5 * torch.randn(121, 6, generator=g)).to(DEV)
    return x, wqkv, bqkv, table


def run(fn, x, wqkv, bqkv, table, att, shift):
    B, H, W, C = x.shape
    _lib.check(fn(_lib.ptr(x), _lib.ptr(wqkv), _lib.ptr(bqkv), _lib.ptr(table), _lib.ptr(att), B, H, W, C, shift, _lib.stream_ptr()))


def timeit(fn, iters=10, warm=3):
    """mean GPU time of `iters` back-to-back calls (the stream-ordered pool is only trimmed at a synchronize, so the
    cudaMallocAsync calls inside the debug entry points stay cheap); the operands (>= 44 MB each way, 354 MB at H = 240) do not
    fit in L2 together with the previous call's output"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    out = {}
    bad = 0
    quick = "--quick" in sys.argv
    for (B, H, W, C, shift) in [] if quick else [(1, 6, 6, 192, 0), (1, 12, 12, 192, 0), (1, 12, 12, 192, 3), (2, 18, 24, 192, 3), (1, 6, 6, 192, 3),
                                (1, 12, 12, 96, 0), (2, 24, 18, 96, 3), (3, 48, 48, 192, 3), (2, 60, 60, 96, 3), (5, 30, 30, 192, 0),
                                (16, 60, 60, 192, 3), (16, 120, 120, 96, 3)]:
        x, wqkv, bqkv, table = make(B, H, W, C, B * H + W + C + shift)
        a0 = torch.full((B, H, W, C), 7.0, dtype=torch.float16, device=DEV)
        a1 = torch.full((B, H, W, C), 7.0, dtype=torch.float16, device=DEV)
        run(L.nb200_swin_attn_fused_f16, x, wqkv, bqkv, table, a0, shift)
        run(L.nb200_swin_attn_tc_f16, x, wqkv, bqkv, table, a1, shift)
        torch.cuda.synchronize()
        d = (a0.float() - a1.float()).abs()
        nbad = int((d > 8e-3).sum().item())
        untouched = int((a1 == 7.0).sum().item())
        print(f"B={B} H={H} W={W} C={C} shift={shift}: max diff {d.max().item():.3e} mean {d.mean().item():.3e} >8e-3: {nbad} untouched: {untouched}",
              flush=True)
        if nbad:
            bad += 1
            idx = (d > 8e-3).nonzero()[:8].tolist()
            print("   first bad (b, y, x, c):", idx)
            bb, yy, xx = idx[0][0], idx[0][1], idx[0][2]
            print("   old:", a0[bb, yy, xx, :8].tolist(), "\n   new:", a1[bb, yy, xx, :8].tolist())
        out[f"diff_B{B}_H{H}_W{W}_C{C}_s{shift}"] = dict(max=d.max().item(), mean=d.mean().item(), bad=nbad)
    if bad:
        print("MISMATCH in", bad, "shapes")
    for (B, H, C) in [(16, 240, 192)] if quick else [(16, 240, 192), (16, 240, 96), (16, 120, 192), (16, 60, 192)]:
        x, wqkv, bqkv, table = make(B, H, H, C, 2)
        att = torch.empty(B, H, H, C, dtype=torch.float16, device=DEV)
        t_old = timeit(lambda: run(L.nb200_swin_attn_fused_f16, x, wqkv, bqkv, table, att, 3))
        t_new = timeit(lambda: run(L.nb200_swin_attn_tc_f16, x, wqkv, bqkv, table, att, 3))
        T = B * H * H
        flop = 2.0 * T * C * 3 * C + 4.0 * T * 36 * C
        print(f"attn B={B} H={H} C={C}: mma.sync warps {t_old:.1f} us, tcgen05 {t_new:.1f} us ({flop / t_new / 1e6:.0f} TF/s algorithmic)", flush=True)
        out[f"attn_B{B}_H{H}_C{C}"] = dict(old_us=t_old, tc_us=t_new, tc_tflops=flop / t_new / 1e6)
    if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
        json.dump(out, open(sys.argv[1], "w"), indent=1)
    if "--timeline" in sys.argv:
        C = int(sys.argv[sys.argv.index("--timeline") + 1])
        x, wqkv, bqkv, table = make(16, 240, 240, C, 2)
        att = torch.empty_like(x)
        run(L.nb200_swin_attn_tc_f16, x, wqkv, bqkv, table, att, 3)
        torch.cuda.synchronize()
        tl = torch.zeros(8 * 2048, dtype=torch.int64, device=DEV)
        _lib.check(L.nb200_debug_timeline(_lib.ptr(tl)))
        run(L.nb200_swin_attn_tc_f16, x, wqkv, bqkv, table, att, 3)
        torch.cuda.synchronize()
        _lib.check(L.nb200_debug_timeline(None))
        ev = tl.cpu().tolist()
        NAMES = {1: "G chunk", 2: "S", 3: "PV", 10: "E wait d", 11: "E d ok", 12: "E qk_full sent", 13: "E v_full sent", 14: "E q/k in regs",
                 15: "E q,k stored", 16: "E d_empty sent", 17: "E v_empty ok", 18: "E v stored", 20: "sm wait S", 21: "sm S ok", 22: "sm S in regs",
                 23: "sm math done", 24: "sm O-epi done", 25: "sm p_full sent", 26: "sm p_empty ok", 27: "sm P stored", 28: "sm o_full ok",
                 29: "sm O in regs", 60: "mask v|p|qk|sfree|dempty|x|w =", 40: "X wait-empty", 41: "X issue", 50: "W wait-empty", 51: "W issue"}
        TRACK = ["mma", "E0", "sm0", "sm1", "prodX", "prodW"]
        rows = []
        for tr in range(6):
            for i in range(2048):
                v = ev[tr * 2048 + i] & ((1 << 64) - 1)
                if v == 0:
                    break
                rows.append((v & 0xffffffffff, tr, (v >> 56) & 0xff, (v >> 40) & 0xffff))
        rows.sort()
        starts = [t for (t, tr, tag, aux) in rows if tr == 4 and tag == 41]
        print(f"C={C}: {len(starts)} tiles seen; cycles per tile (X issue):", [starts[i + 1] - starts[i] for i in range(min(len(starts) - 1, 12))])
        if len(starts) > 6:
            t0, t1 = starts[3], starts[5]
            for (t, tr, tag, aux) in rows:
                if t0 <= t <= t1:
                    print(f"{t - t0:8d}  {TRACK[tr]:6s} {NAMES.get(tag, tag)} [{aux if tag != 60 else format(aux & 127, '07b')[::-1] + ' kc' + str(aux >> 8)}]")


main()
