import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from tests.util import load_golden
from nunif_b200.iw3 import batch_preprocess
from oracle import frames as ofr
g = load_golden("frames")
x = torch.from_numpy(g["x"]).cuda()
got = batch_preprocess(x, lower_bound=126).cpu().numpy()
want = g["prep_126"]
d = np.abs(got - want)
print("max", d.max(), "mean", d.mean())
idx = np.unravel_index(np.argsort(d.ravel())[-8:], d.shape)
for k in range(8):
    i = tuple(int(a[k]) for a in idx)
    print(i, got[i], want[i], d[i])
print("err by column mod:", [float(d[..., c::8].mean()) for c in range(8)])
print("err by row:", [float(d[:, :, r].mean()) for r in range(0, 126, 16)])
print("err by channel:", [float(d[:, c].mean()) for c in range(3)])
