"""Developer diagnostic (GPU): fused kernels vs goldens; prints mismatch counts and error levels."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
import e2e_slam_b200 as e2e
dev = torch.device("cuda:0")
def same(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
def relmax(a, b): return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
for f in sorted(glob.glob(os.path.join(ROOT, "tests/golden/*.npz"))):
    g = np.load(f); pad = str(g["padding_mode"]); mask = bool(g["use_mask"])
    t = lambda n: torch.from_numpy(g[n]).to(dev)
    colors = t("colors").requires_grad_(True)
    depth = t("depth").requires_grad_(True); T = t("T").requires_grad_(True)
    src = colors[:, 0].permute(0, 3, 1, 2); tgt = colors[:, 1].permute(0, 3, 1, 2)
    lm, syn, valid, pix = e2e.warp_photometric(depth, t("inv_K"), t("K"), T, src, tgt, pad, mask, need_outputs=True)
    lm.mean().backward()
    out = [os.path.basename(f)]
    for k, v in (("pix", pix), ("valid", valid), ("syn", syn), ("loss_map", lm)):
        out.append(f"{k}:{same(v.detach().cpu().numpy(), g[k])}")
    gs = colors.grad[:, 0].cpu().numpy()
    for k, v in (("g_depth", depth.grad.cpu().numpy()), ("g_src", gs), ("g_T", T.grad.cpu().numpy())):
        out.append(f"{k}: vs32 {relmax(v, g[k]):.2e} vs64 {relmax(v, g[k+'_f64']):.2e} (ref32 vs64 {relmax(g[k], g[k+'_f64']):.2e})")
    # lean path
    depth2 = t("depth").requires_grad_(True)
    l = e2e.warp_photometric_loss(depth2, t("inv_K"), t("K"), t("T"), src.detach(), tgt.detach(), pad, mask)
    l.backward()
    out.append(f"loss {float(l):.8f} ref {float(g['loss']):.8f} rel {abs(float(l)-float(g['loss']))/float(g['loss']):.1e} lean g_depth vs32 {relmax(depth2.grad.cpu().numpy(), g['g_depth']):.2e}")
    print("\n  ".join(out))
print("launches", e2e.ops.launch_count())
