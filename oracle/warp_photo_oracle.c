/*
 * CPU oracle, plain C -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Explicit-operation-order restatement of the reference's forward path
 *   backproject -> project -> normalise -> valid mask -> bilinear sample -> mask -> SSIM + L1
 * in IEEE-754 binary32, one rounding per written operation.  fmaf() appears exactly where the
 * reference's CPU execution uses a fused multiply-add (BLAS k-loops of torch.matmul; the
 * compiler-contracted multiply-adds inside ATen's vectorised grid sampler); everything else is a
 * separately rounded +, -, *, /.  Compile with -ffp-contract=off so the compiler adds none.
 *
 * Parity: pinned bit-for-bit against tests/golden/*.npz (outputs of the real reference, see
 * tools/make_golden.py) by tests/test_oracle_golden.py.  The CUDA kernels mirror this file's
 * operation order and are tested for bit equality against it at sizes the goldens do not cover.
 *
 * Reference lines (relative to the reference repo root):
 *   depth_estimation/view_synthesis.py:17-31  pixel grid [x, y, 1], j = y*W + x
 *   depth_estimation/view_synthesis.py:36-38  X = (inv_K[:3,:3] @ pix) * depth
 *   depth_estimation/view_synthesis.py:57-60  P = (K@T)[:3]; c = P @ [X;1]; uv = c[:2] / (c[2] + eps)
 *   depth_estimation/view_synthesis.py:66-71  /(W-1), /(H-1), (p-0.5)*2, valid = max|p| <= 1
 *   train_depth.py:587-590                    F.grid_sample(..., align_corners=False)
 *   train_depth.py:713-718                    prediction*mask, target*mask
 *   loss/losses.py:23-37                      SSIM (reflect pad 1, five 3x3 mean pools)
 *   loss/losses.py:111-115                    0.85*mean_c(ssim) + 0.15*mean_c|t-p|
 *
 * Layouts: depth (B,1,H,W); inv_K, K, T (B,4,4) row-major; src/tgt channels-last (B,H,W,3), which
 * is the memory the reference's NCHW *views* alias (train_depth.py:451-453).
 * Outputs (each may be NULL): pix (B,H,W,2); valid (B,H,W); syn, ssim (B,3,H,W); loss_map (B,H,W).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#define C1F 1.0e-4f   /* 0.01**2 evaluated in double by Python, then rounded to float by torch */
#define C2F 9.0e-4f   /* 0.03**2 */

static int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

/* K @ T rows 0..2: at::bmm's small-matrix path accumulates k = 0..3 in order; products are
 * exactly representable or not, either way a sequential multiply-add (unfused) reproduces it. */
static void compose_P(const float *K, const float *T, float *P)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) {
            float acc = 0.0f;
            for (int k = 0; k < 4; k++)
#ifdef ORACLE_P_FMA
                acc = fmaf(K[i * 4 + k], T[k * 4 + j], acc);
#else
                acc = acc + K[i * 4 + k] * T[k * 4 + j];
#endif
            P[i * 4 + j] = acc;
        }
}

int e2e_oracle_warp_photo_fwd(const float *depth, const float *inv_K, const float *K, const float *T,
                              const float *src, const float *tgt, int B, int H, int W,
                              int padding_border, int use_mask, float eps,
                              float *pix, float *valid, float *syn, float *ssim, float *loss_map)
{
    const size_t HW = (size_t)H * W;
    float *xs = (float *)malloc(sizeof(float) * 3 * HW);
    float *ys = (float *)malloc(sizeof(float) * 3 * HW);
    if (!xs || !ys) { free(xs); free(ys); return 1; }
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float half_w = (float)W / 2, half_h = (float)H / 2;

    for (int b = 0; b < B; b++) {
        const float *ik = inv_K + b * 16;
        float P[12];
        compose_P(K + b * 16, T + b * 16, P);
        const float *sb = src + (size_t)b * HW * 3, *tb = tgt + (size_t)b * HW * 3;

#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const size_t j = (size_t)y * W + x;
                const float fx = (float)x, fy = (float)y, d = depth[b * HW + j];
                float X[3], c[3];
                for (int i = 0; i < 3; i++) {                       /* sgemm k-loop, then * depth */
                    float a = ik[i * 4 + 0] * fx;
                    a = fmaf(ik[i * 4 + 1], fy, a);
                    a = fmaf(ik[i * 4 + 2], 1.0f, a);
                    X[i] = d * a;
                }
                for (int i = 0; i < 3; i++) {                       /* sgemm k-loop over [X;1] */
                    float a = P[i * 4 + 0] * X[0];
                    a = fmaf(P[i * 4 + 1], X[1], a);
                    a = fmaf(P[i * 4 + 2], X[2], a);
                    a = fmaf(P[i * 4 + 3], 1.0f, a);
                    c[i] = a;
                }
                const float z = c[2] + eps;
                float gx = (c[0] / z) / wm1, gy = (c[1] / z) / hm1;
                gx = (gx - 0.5f) * 2.0f;
                gy = (gy - 0.5f) * 2.0f;
                const float v = (fabsf(gx) <= 1.0f && fabsf(gy) <= 1.0f) ? 1.0f : 0.0f;   /* NaN -> 0 */
                if (pix) { pix[(b * HW + j) * 2] = gx; pix[(b * HW + j) * 2 + 1] = gy; }
                if (valid) valid[b * HW + j] = v;

                /* grid_sample, bilinear, align_corners=False (ATen cpu/GridSamplerKernel.cpp) */
                float ix = fmaf(gx + 1.0f, half_w, -0.5f), iy = fmaf(gy + 1.0f, half_h, -0.5f);
                if (padding_border) {                               /* NaN clamps to 0 */
                    ix = fminf(wm1, fmaxf(0.0f, ix));
                    iy = fminf(hm1, fmaxf(0.0f, iy));
                }
                const float xw = floorf(ix), yn = floorf(iy);
                const float w = ix - xw, e = 1.0f - w, n = iy - yn, s = 1.0f - n;
                const float nw = s * e, ne = s * w, sw = n * e, se = n * w;
                /* float -> int with saturation, so wildly out-of-range coordinates stay out of range */
                const float xc = fminf(fmaxf(xw, -2.0f), (float)W + 1.0f), yc = fminf(fmaxf(yn, -2.0f), (float)H + 1.0f);
                const int x0 = (xw == xw) ? (int)xc : -2, y0 = (yn == yn) ? (int)yc : -2;
                const int x1 = x0 + 1, y1 = y0 + 1;
                const int inx0 = x0 >= 0 && x0 < W, inx1 = x1 >= 0 && x1 < W;
                const int iny0 = y0 >= 0 && y0 < H, iny1 = y1 >= 0 && y1 < H;
                for (int ch = 0; ch < 3; ch++) {
                    const float a = (inx0 && iny0) ? sb[((size_t)y0 * W + x0) * 3 + ch] : 0.0f;
                    const float bb = (inx1 && iny0) ? sb[((size_t)y0 * W + x1) * 3 + ch] : 0.0f;
                    const float cc = (inx0 && iny1) ? sb[((size_t)y1 * W + x0) * 3 + ch] : 0.0f;
                    const float dd = (inx1 && iny1) ? sb[((size_t)y1 * W + x1) * 3 + ch] : 0.0f;
                    float o = a * nw;
                    o = fmaf(bb, ne, o);
                    o = fmaf(cc, sw, o);
                    o = fmaf(dd, se, o);
                    if (syn) syn[((size_t)b * 3 + ch) * HW + j] = o;
                    const float t = tb[j * 3 + ch];
                    xs[ch * HW + j] = use_mask ? o * v : o;
                    ys[ch * HW + j] = use_mask ? t * v : t;
                }
            }

        if (!ssim && !loss_map) continue;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const size_t j = (size_t)y * W + x;
                float sv[3], lv[3];
                for (int ch = 0; ch < 3; ch++) {
                    const float *px = xs + ch * HW, *py = ys + ch * HW;
                    float sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;   /* avg_pool2d: kh outer, kw inner */
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) {
                            const size_t q = (size_t)reflect1(y + dy, H) * W + reflect1(x + dx, W);
                            const float a = px[q], bq = py[q];
                            sx += a; sy += bq; sxx += a * a; syy += bq * bq; sxy += a * bq;
                        }
                    const float mx = sx / 9.0f, my = sy / 9.0f;
                    const float vx = sxx / 9.0f - mx * mx, vy = syy / 9.0f - my * my, vxy = sxy / 9.0f - mx * my;
                    const float nn = (2.0f * mx * my + C1F) * (2.0f * vxy + C2F);
                    const float dn = (mx * mx + my * my + C1F) * (vx + vy + C2F);
                    float q = (1.0f - nn / dn) / 2.0f;
                    q = q < 0.0f ? 0.0f : (q > 1.0f ? 1.0f : q);      /* NaN propagates like torch.clamp */
                    sv[ch] = q;
                    lv[ch] = fabsf(py[j] - px[j]);
                    if (ssim) ssim[((size_t)b * 3 + ch) * HW + j] = q;
                }
                if (loss_map) {
                    const float sm = ((sv[0] + sv[1]) + sv[2]) / 3.0f, lm = ((lv[0] + lv[1]) + lv[2]) / 3.0f;
                    loss_map[b * HW + j] = 0.85f * sm + 0.15f * lm;
                }
            }
    }
    free(xs);
    free(ys);
    return 0;
}
