"""Autograd map path (what patch.fuse() gives the reference's scripts): forward with the tensors the scripts keep +
`.mean(1, keepdim=True).mean()` + backward.  Run with E2E_SPECULATE=0 for forward kernel + unconditional backward kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
import e2e_slam_b200 as e2e
from e2e_slam_b200.synthetic import make_pairs
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, W = 480, 640
dev = torch.device("cuda:0")
chunks = [make_pairs(min(32, P - s), H, W, "icl", seed=s, device=dev) for s in range(0, P, 32)]
d = {k: torch.cat([c[k] for c in chunks]) for k in chunks[0]}
src0, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)


def step(weighted=False):
    depth = d["depth"].detach().requires_grad_(True)
    src = src0.detach().requires_grad_(True)
    T = d["T"].detach().requires_grad_(True)
    lm, syn, valid, pix = e2e.warp_photometric(depth, d["inv_K"], d["K"], T, src, tgt, "border", True, need_outputs=True)
    m = lm.mean(1, keepdim=True)
    loss = (m * valid).mean() if weighted else m.mean()
    loss.backward()
    return loss, depth.grad, src.grad, T.grad


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print(f"E2E_SPECULATE={os.environ.get('E2E_SPECULATE', '1')} pairs={P}: uniform upstream {timeit(step):.3f} ms, weighted upstream {timeit(lambda: step(True)):.3f} ms per step")
