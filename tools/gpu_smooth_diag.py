import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200"))
from e2e_slam_b200.losses import smoothness_loss
from e2e_slam_b200.synthetic import make_pairs
from oracle import torch_oracle
B, H, W = 1, 300, 300
d = make_pairs(B, H, W, "tum", seed=9)
disp = (1.0 / (d["depth"] + 0.3)); img = d["colors"][:, 1].permute(0, 3, 1, 2)
do = disp.clone().requires_grad_(True); lo = torch_oracle.smoothness(do, img); lo.backward()
dg = disp.cuda().requires_grad_(True); lg = smoothness_loss(dg, img.cuda()); lg.backward()
a = dg.grad.cpu().numpy().astype(np.float64)[0, 0]; r = do.grad.numpy().astype(np.float64)[0, 0]
diff = a - r; bad = np.abs(diff) > 1e-3 * np.abs(r).max()
print("bad", int(bad.sum()), "of", bad.size, "max|r|", np.abs(r).max())
ys, xs = np.nonzero(bad)
print("rows hist", np.bincount(ys, minlength=H)[:40], "...")
print("cols hist", np.bincount(xs, minlength=W)[:40], "...")
flat = ys * W + xs
print("flat idx first 30", flat[:30])
print("flat idx mod 256 hist nonzero", np.nonzero(np.bincount(flat % 256, minlength=256))[0][:50])
print("diff/maxr first 10", (diff[bad] / np.abs(r).max())[:10])
