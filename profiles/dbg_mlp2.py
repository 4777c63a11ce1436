import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib
DEV = "cuda:0"; L = _lib.lib()
C = 192
for T in (57600, 230400, 921600):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(T, C, generator=g).half().to(DEV)
    w1 = (torch.randn(2 * C, C, generator=g) / C ** 0.5).half().to(DEV)
    w2 = (torch.randn(C, 2 * C, generator=g) / (2 * C) ** 0.5).half().to(DEV)
    b1, b2 = torch.zeros(2 * C, device=DEV), torch.zeros(C, device=DEV)
    def run():
        _lib.check(L.nb200_swin_mlp_fused_f16(_lib.ptr(x), None, T, C, None, None, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.stream_ptr()))
    for _ in range(2): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(round(a.elapsed_time(b) * 1e3, 1))
    print("T", T, "us per call:", ts, flush=True)
