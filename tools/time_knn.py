"""kNN timing: brute force vs grid on a surface-like map (python tools/time_knn.py P1 P2)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from e2e_slam_b200._lib import check, lib, ptr, stream_ptr
P1 = int(sys.argv[1]) if len(sys.argv) > 1 else 307200
P2 = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
g = torch.Generator(device="cuda").manual_seed(5)
uv = torch.rand(P2, 2, generator=g, device="cuda") * 8 - 4
r = torch.stack([uv[:, 0], uv[:, 1], 3.0 + 0.4 * torch.sin(uv[:, 0]) * torch.cos(1.3 * uv[:, 1])], 1).contiguous()
OFF = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0      # queries pushed off the surface by N(0, OFF) along z
q = (r[torch.randint(0, P2, (P1,), device="cuda", generator=g)] + 0.01 * (torch.rand(P1, 3, generator=g, device="cuda") - 0.5))
q[:, 2] += OFF * torch.randn(P1, generator=g, device="cuda")
q = q.contiguous()
d2 = torch.empty(P1, device="cuda"); idx = torch.empty(P1, dtype=torch.int64, device="cuda")
nws = lib().e2e_knn1_grid_workspace_bytes(P2); ws = torch.empty(nws, dtype=torch.uint8, device="cuda")


def t(fn, n):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


tg = t(lambda: check(lib().e2e_knn1_grid_fwd(ptr(q), None, ptr(r), P1, P2, ptr(d2), ptr(idx), ptr(ws), nws, stream_ptr()), "g"), 5)
ig = idx.clone()
if len(sys.argv) > 4:
    print(f"P1={P1} P2={P2} off={OFF}: grid {tg:.3f} ms (mean nn distance {float(d2.sqrt().mean()):.4f})")
else:
    tb = t(lambda: check(lib().e2e_knn1_fwd(ptr(q), None, ptr(r), P1, P2, ptr(d2), ptr(idx), stream_ptr()), "b"), 1)
    print(f"P1={P1} P2={P2} off={OFF}: grid {tg:.3f} ms, brute force {tb:.1f} ms, equal={bool(torch.equal(ig, idx))}")
