"""Closed-form backward of the warp + SSIM/L1 path in numpy float64 -- TEST INFRASTRUCTURE ONLY.

Purpose: write down, once, the per-pixel gradient formulas that the CUDA backward kernel
implements and check them against torch.autograd applied to the forward
restatement (oracle/torch_oracle.py) in float64 -- see tests/test_oracle_golden.py.  It is a
second, independent route to the same gradients; nothing in the product imports it.

Forward lines differentiated: depth_estimation/view_synthesis.py:34-40, 54-71;
F.grid_sample(align_corners=False) at train_depth.py:587-590; train_depth.py:713-718;
loss/losses.py:23-37, 111-115.
"""
import numpy as np

C1 = 0.01 ** 2
C2 = 0.03 ** 2


def _reflect(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def backward(depth, inv_K, K, T, src_cl, tgt_cl, padding_mode="border", use_mask=True, eps=1e-7,
             grad_loss_map=None):
    """Returns dict(g_depth (B,1,H,W), g_src (B,H,W,3), g_P (B,3,4), g_T (B,4,4), loss).

    grad_loss_map: upstream dL/d loss_map (B,1,H,W); default = 1/(B*H*W) (i.e. L = loss_map.mean())."""
    f8 = lambda a: np.asarray(a, dtype=np.float64)
    depth, inv_K, K, T, src_cl, tgt_cl = map(f8, (depth, inv_K, K, T, src_cl, tgt_cl))
    B, _, H, W = depth.shape
    if grad_loss_map is None:
        grad_loss_map = np.full((B, 1, H, W), 1.0 / (B * H * W))
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    g_depth = np.zeros((B, 1, H, W))
    g_src = np.zeros((B, H, W, 3))
    g_P = np.zeros((B, 3, 4))
    g_T = np.zeros((B, 4, 4))
    total = 0.0
    for b in range(B):
        ik = inv_K[b, :3, :3]
        P = (K[b] @ T[b])[:3]
        A, t = P[:, :3], P[:, 3]
        r = np.stack([ik[i, 0] * xs + ik[i, 1] * ys + ik[i, 2] for i in range(3)], 0)      # (3,H,W)
        X = depth[b, 0][None] * r
        c = np.einsum("ij,jhw->ihw", A, X) + t[:, None, None]
        z = c[2] + eps
        u, v = c[0] / z, c[1] / z
        gx, gy = (u / (W - 1) - 0.5) * 2, (v / (H - 1) - 0.5) * 2
        valid = ((np.abs(gx) <= 1) & (np.abs(gy) <= 1)).astype(np.float64)
        ix, iy = (gx + 1) * (W / 2) - 0.5, (gy + 1) * (H / 2) - 0.5
        mx_ = np.ones_like(ix)
        my_ = np.ones_like(iy)
        if padding_mode == "border":
            # ATen clip_coordinates_set_grad: zero gradient when ix <= 0 or ix >= size-1
            mx_ = ((ix > 0) & (ix < W - 1)).astype(np.float64)
            my_ = ((iy > 0) & (iy < H - 1)).astype(np.float64)
            ix, iy = np.clip(ix, 0, W - 1), np.clip(iy, 0, H - 1)
        x0, y0 = np.floor(ix), np.floor(iy)
        wx, wy = ix - x0, iy - y0
        x0, y0 = x0.astype(np.int64), y0.astype(np.int64)

        def tap(yy, xx):
            inb = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            val = src_cl[b, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * inb[..., None]
            return val, inb

        v00, i00 = tap(y0, x0)
        v10, i10 = tap(y0, x0 + 1)
        v01, i01 = tap(y0 + 1, x0)
        v11, i11 = tap(y0 + 1, x0 + 1)
        w00, w10, w01, w11 = (1 - wx) * (1 - wy), wx * (1 - wy), (1 - wx) * wy, wx * wy
        syn = v00 * w00[..., None] + v10 * w10[..., None] + v01 * w01[..., None] + v11 * w11[..., None]
        m = valid if use_mask else np.ones_like(valid)
        xm = syn * m[..., None]                       # prediction  (H,W,3)
        ym = tgt_cl[b] * m[..., None]                 # target

        # ---- SSIM statistics over the reflect-padded 3x3 window ---------------------------------
        ry = _reflect(np.arange(-1, H + 1), H)
        rx = _reflect(np.arange(-1, W + 1), W)
        xp, yp = xm[ry][:, rx], ym[ry][:, rx]         # (H+2, W+2, 3)

        def box(a):
            return sum(a[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)) / 9.0

        mux, muy = box(xp), box(yp)
        sxx, syy, sxy = box(xp * xp), box(yp * yp), box(xp * yp)
        A1 = 2 * mux * muy + C1
        A2 = 2 * (sxy - mux * muy) + C2
        B1 = mux ** 2 + muy ** 2 + C1
        B2 = (sxx - mux ** 2) + (syy - muy ** 2) + C2
        n, dn = A1 * A2, B1 * B2
        Q = n / dn
        s_raw = (1 - Q) / 2
        ssim = np.clip(s_raw, 0, 1)
        lmap = 0.85 * ssim.mean(-1) + 0.15 * np.abs(ym - xm).mean(-1)
        total += float((lmap * grad_loss_map[b, 0]).sum())

        # ---- backward: SSIM + L1 -> prediction ---------------------------------------------------
        g = grad_loss_map[b, 0][..., None]             # dL/d loss_map, broadcast over channels
        g_s = 0.85 / 3.0 * g * ((s_raw >= 0) & (s_raw <= 1))      # clamp passes gradient on [0,1] inclusive
        dQ_dmux = (2 * muy * (A2 - A1) * dn - n * 2 * mux * (B2 - B1)) / dn ** 2
        dQ_dsxx = -n * B1 / dn ** 2
        dQ_dsxy = 2 * A1 / dn
        Ga, Gb, Gc = -0.5 * g_s * dQ_dmux, -0.5 * g_s * dQ_dsxx, -0.5 * g_s * dQ_dsxy

        def box_t(G):
            """adjoint of reflect-pad(1) + 3x3 mean: spread G/9 onto the padded image, fold the pad back."""
            pad = np.zeros((H + 2, W + 2, 3))
            for dy in range(3):
                for dx in range(3):
                    pad[dy:dy + H, dx:dx + W] += G / 9.0
            out = np.zeros((H, W, 3))
            np.add.at(out, (ry[:, None], rx[None, :]), pad)
            return out

        g_x = box_t(Ga) + 2 * xm * box_t(Gb) + ym * box_t(Gc) + 0.15 / 3.0 * g * np.sign(xm - ym)
        g_syn = g_x * m[..., None]

        # ---- backward: sampler -> source image and sample position -------------------------------
        for (yy, xx, inb, w) in ((y0, x0, i00, w00), (y0, x0 + 1, i10, w10), (y0 + 1, x0, i01, w01), (y0 + 1, x0 + 1, i11, w11)):
            contrib = g_syn * (w * inb)[..., None]
            np.add.at(g_src[b], (np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)), contrib)
        gix = (g_syn * ((v10 - v00) * (1 - wy)[..., None] + (v11 - v01) * wy[..., None])).sum(-1) * mx_
        giy = (g_syn * ((v01 - v00) * (1 - wx)[..., None] + (v11 - v10) * wx[..., None])).sum(-1) * my_
        gu, gv = gix * W / (W - 1), giy * H / (H - 1)
        gc0, gc1 = gu / z, gv / z
        gc2 = -(gu * c[0] + gv * c[1]) / z ** 2
        gc = np.stack([gc0, gc1, gc2], 0)              # (3,H,W)
        q = np.einsum("ij,jhw->ihw", A, r)
        g_depth[b, 0] = (gc * q).sum(0)
        X1 = np.concatenate([X, np.ones((1, H, W))], 0)
        g_P[b] = np.einsum("ihw,jhw->ij", gc, X1)
        g_T[b] = K[b, :3, :].T @ g_P[b]
    return dict(g_depth=g_depth, g_src=g_src, g_P=g_P, g_T=g_T, loss=total)
