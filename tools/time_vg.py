"""Event-timed comparison of the single-pass value+gradient kernel with the forward + backward pair.
    python tools/time_vg.py [pairs] [iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")); sys.path.insert(0, ROOT)
from e2e_slam_b200 import ops
from e2e_slam_b200.synthetic import make_pairs
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H = int(sys.argv[3]) if len(sys.argv) > 3 else 480
W = int(sys.argv[4]) if len(sys.argv) > 4 else 640
dev = torch.device("cuda:0")
chunks = [make_pairs(min(32, P - s), H, W, "icl", seed=s, device=dev) for s in range(0, P, 32)]
d = {k: torch.cat([c[k] for c in chunks]) for k in chunks[0]}
src, tgt = d["colors"][:, 0].permute(0, 3, 1, 2), d["colors"][:, 1].permute(0, 3, 1, 2)
plan = ops.WarpPhotoPlan(P, H, W, dev)
a = (d["depth"], d["inv_K"], d["K"], d["T"], src, tgt)


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_f = timeit(lambda: plan.forward(*a), iters)
lf = float(plan.loss)
t_b = timeit(lambda: plan.backward(*a), iters)
gd_b, gs_b, gp_b = plan.grad_depth.clone(), plan.grad_src.clone(), plan.grad_P.clone()
t_v = timeit(lambda: plan.value_and_grad(*a), iters)
lv = float(plan.loss)
gd_v, gs_v, gp_v = plan.grad_depth.clone(), plan.grad_src.clone(), plan.grad_P.clone()
t_z = timeit(lambda: plan.grad_src.zero_(), iters)
npx = P * H * W
rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
print(f"pairs={P} fwd {t_f:.3f} ms  bwd(+zero) {t_b:.3f} ms  vg(+zero) {t_v:.3f} ms  zero {t_z:.3f} ms")
print(f"vg: {npx / t_v / 1e6:.2f} Gpx/s, {72 * npx / t_v / 1e6:.0f} GB/s algorithmic (72 B/px)")
print(f"loss fwd {lf:.8f} vg {lv:.8f}  rel diffs vs bwd kernel: depth {rel(gd_v, gd_b):.2e} src {rel(gs_v, gs_b):.2e} P {rel(gp_v, gp_b):.2e}")
