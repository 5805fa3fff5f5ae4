"""Per-phase times of the whole-sequence fusion kernel (library built with E2E_NVCC_FLAGS=-DE2E_SEQ_TIMING)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "end-to-end-self-supervised-slam_b200"))
import torch  # noqa: E402
from e2e_slam_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402
from e2e_slam_b200.synthetic import room_sequence  # noqa: E402

L, H, W = 60, 480, 640
dev = torch.device("cuda", 0)
depth, rgb, K, poses = room_sequence(L, H, W, device=dev)
cap = L * H * W
z = dict(dtype=torch.float32, device=dev)
a256 = lambda n: (n + 255) // 256 * 256
hw = H * W
off_counts = 3 * a256(cap * 16) + 2 * a256(hw * 16) * 3
for rep in range(3):
    pts, nrm, col, cc = torch.empty(cap, 3, **z), torch.empty(cap, 3, **z), torch.empty(cap, 3, **z), torch.empty(cap, **z)
    n = torch.zeros(2, dtype=torch.int64, device=dev)
    nws = lib().e2e_fusion_sequence_workspace_bytes(H, W, cap)
    ws = torch.zeros(nws, dtype=torch.uint8, device=dev)
    check(lib().e2e_fusion_sequence(ptr(depth), ptr(rgb), ptr(K), ptr(poses), L, H, W, ctypes.c_float(0.6), ctypes.c_float(0.05),
                                    ctypes.c_float(0.9396926), ptr(pts), ptr(nrm), ptr(col), ptr(cc), ptr(n), 0, cap, ptr(ws), nws,
                                    stream_ptr()), "seq")
    torch.cuda.synchronize()
t = ws[off_counts:off_counts + 8 * 4096].view(torch.int64)[2048:2048 + 2 + 3 * L].cpu().double()
d = (t[1:] - t[:-1]) / 1e3
print("maps0 %.1f us" % d[0])
p1, p2, p3 = d[1::3], d[2::3], d[3::3]
for name, v in (("P1", p1), ("P2", p2), ("P3", p3)):
    print(name, "mean %.1f  first10 %.1f  last10 %.1f  max %.1f us" % (v.mean(), v[:10].mean(), v[-10:].mean(), v.max()))
q = ws[off_counts:off_counts + 8 * 4096].view(torch.int64)[2048 + 256:2048 + 256 + 4 * L].cpu().double().view(L, 4)
p3_start = t[3::3][:L]          # stamp after the P2 barrier
p3_end = t[4::3][:L]
seg = torch.stack([q[:, 0] - p3_start, q[:, 1] - q[:, 0], q[:, 2] - q[:, 1], q[:, 3] - q[:, 2], p3_end - q[:, 3]], 1) / 1e3
print("P3 (CTA 0): merge %.1f | scan+publish %.1f | maps(s+1) %.1f | look-back %.1f | append+barrier %.1f us" % tuple(seg[:-1].mean(0)))
p1_start, p2_start = torch.cat([t[1:2], t[4::3][:L - 1]]), t[2::3][:L]
for which, name in enumerate(("CTA 0", "CTA mid", "CTA last")):
    o = ws[off_counts:off_counts + 8 * 4096].view(torch.int64)[2048 + 512 + which * 128:2048 + 512 + which * 128 + 2 * L].cpu().double().view(L, 2)
    print("%s own work: P1 %.1f of %.1f us, P2 %.1f of %.1f us" % (name, ((o[:, 0] - p1_start) / 1e3).mean(), p1.mean(), ((o[:, 1] - p2_start) / 1e3).mean(), p2.mean()))
bb = ws[off_counts:off_counts + 8 * 4096].view(torch.int64)[2048 + 1024:2048 + 1024 + 17].cpu().double()
print("empty barriers (us):", " ".join("%.2f" % x for x in ((bb[1:] - bb[:-1]) / 1e3).tolist()))
print("total %.1f us" % ((t[-1] - t[0]) / 1e3), "N =", int(n[0]))
