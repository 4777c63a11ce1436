import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nunif_b200 import _lib
DEV = "cuda:0"
B, H, W, C, shift = 1, 12, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 192, 0
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, C, generator=g).half().to(DEV)
wqkv = (torch.randn(3 * C, C, generator=g) / C ** 0.5).half().to(DEV)
bqkv = torch.zeros(3 * C, device=DEV)
table = torch.zeros(121, 6, device=DEV)
att = torch.zeros(B, H, W, C, dtype=torch.float16, device=DEV)
_lib.check(_lib.lib().nb200_swin_attn_fused_f16(_lib.ptr(x), _lib.ptr(wqkv), _lib.ptr(bqkv), _lib.ptr(table), _lib.ptr(att), B, H, W, C, shift, _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok", float(att.float().abs().mean()))
