"""Achieved bandwidth of the granular (tier (i)) ops at B x 480 x 640: backproject, project3d, grid_sample, SSIM, photometric_loss."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200 import losses, view_synthesis
from e2e_slam_b200.synthetic import make_pairs
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 480, 640
dev = torch.device("cuda:0")
chunks = [make_pairs(min(32, B - s), H, W, "icl", seed=s, device=dev) for s in range(0, B, 32)]
d = {k: torch.cat([c[k] for c in chunks]) for k in chunks[0]}
src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
bp, pr, ssim = view_synthesis.BackprojectDepth(B, H, W), view_synthesis.Project3D(B, H, W), losses.SSIM()


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def chain(grad):
    depth = d["depth"].detach().requires_grad_(grad)
    s = src.detach().requires_grad_(grad)
    pts = bp(depth, d["inv_K"])
    pix, valid = pr(pts, d["K"], d["T"], False)
    syn = view_synthesis.grid_sample(s, pix, padding_mode="border", align_corners=False)
    lm = losses.photometric_loss(ssim, syn * valid, tgt * valid)
    if grad:
        lm.mean().backward()


npx = B * H * W
with torch.no_grad():
    depth = d["depth"]
    pts = bp(depth, d["inv_K"])
    pix, valid = pr(pts, d["K"], d["T"], False)
    syn = view_synthesis.grid_sample(src, pix, padding_mode="border", align_corners=False)
    for name, fn, bpp in (("backproject fwd", lambda: bp(depth, d["inv_K"]), 4 + 16),
                          ("project3d fwd", lambda: pr(pts, d["K"], d["T"], False), 16 + 8 + 4),
                          ("grid_sample fwd", lambda: view_synthesis.grid_sample(src, pix, padding_mode="border", align_corners=False), 12 + 8 + 12),
                          ("SSIM fwd", lambda: ssim(syn, tgt), 24 + 12),
                          ("photometric_loss fwd", lambda: losses.photometric_loss(ssim, syn, tgt), 24 + 4)):
        t = timeit(fn)
        print(f"{name:24s} {t:7.3f} ms  {bpp * npx / t / 1e6:6.0f} GB/s ({bpp} B/px)")
    t = timeit(lambda: chain(False))
    print(f"{'granular chain fwd':24s} {t:7.3f} ms  {npx / t / 1e6:6.2f} Gpx/s")
t = timeit(lambda: chain(True))
print(f"{'granular chain fwd+bwd':24s} {t:7.3f} ms  {npx / t / 1e6:6.2f} Gpx/s   (the unmodified scripts' call sequence; the fused op does this in one sweep)")


def fb(name, make, bpp):
    def run():
        out, leaves = make()
        out.mean().backward()
    t = timeit(run)
    print(f"{name:24s} {t:7.3f} ms  {bpp * npx / t / 1e6:6.0f} GB/s ({bpp} B/px fwd+bwd incl. autograd mean)")


def mk_ssim():
    a = syn.detach().requires_grad_(True)
    return ssim(a, tgt), (a,)


def mk_photo():
    a = syn.detach().requires_grad_(True)
    return losses.photometric_loss(ssim, a, tgt), (a,)


def mk_gs():
    s = src.detach().requires_grad_(True)
    g = pix.detach().requires_grad_(True)
    return view_synthesis.grid_sample(s, g, padding_mode="border", align_corners=False), (s, g)


def mk_proj():
    p_ = pts.detach().requires_grad_(True)
    return pr(p_, d["K"], d["T"], False)[0], (p_,)


def mk_bp():
    dd = d["depth"].detach().requires_grad_(True)
    return bp(dd, d["inv_K"]), (dd,)


fb("SSIM fwd+bwd", mk_ssim, 36 + 36)
fb("photometric fwd+bwd", mk_photo, 28 + 28)
fb("grid_sample fwd+bwd", mk_gs, 32 + 44)
fb("project3d fwd+bwd", mk_proj, 28 + 24)
fb("backproject fwd+bwd", mk_bp, 20 + 20)
