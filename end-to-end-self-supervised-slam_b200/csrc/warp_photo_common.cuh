// Device helpers shared by the warp + photometric kernels (warp_photo.cu, warp_photo_fused.cu):
// exact-order projection, bilinear sampler set-up, reflection fix-up, 3x3 window sums, SSIM algebra.
#pragma once
#include "common.cuh"

namespace e2e {

#define C1F 1.0e-4f
#define C2F 9.0e-4f

enum { MODE_WARP = 0, MODE_DIRECT = 1 };

struct WPParams {
    // inputs
    const float *depth, *inv_K, *K, *T;
    ImgView src, tgt;          // WARP: source / target image.  DIRECT: x / y.
    int B, C, H, W;            // C = channels of the tensors (DIRECT may be != 3; kernel planes = B*C/CK)
    int border, use_mask, div_exact;
    float eps, wm1, hm1, half_w, half_h;
    float rcpW, rcpH;          // RN(1/(W-1)), RN(1/(H-1))
    // forward outputs (nullable)
    float *syn, *valid, *pix, *ssim, *loss_map, *partial;
    // backward
    const float *g_loss_map, *g_ssim, *g_scalar;
    float g_scale;
    float *g_depth;
    ImgViewW g_src;
    float *gP_partial;
    float *g_x, *g_y;
    const float *skip_flag;    // backward kernels return at once when *skip_flag != 0 (conditional backward)
    int S;                     // streaming kernel: source frames per target (0 / 1 = one).  blockIdx.z = pair * S + source: depth, target and
                               //   intrinsics are indexed by the pair, pose / source image / every output by (pair, source)
    int disp_mode;             // streaming kernel: `depth` holds the network's DISPARITY; depth = (1 / disp) * ratio is formed at the load
    const float *ratio;        //   device scalar of the median scaling (online_adaption.py:295-298), NULL = none
};

// Per-CTA image handle: batch offset applied, 32-bit element strides (host checks they fit).
struct Img32 {
    const float *p;
    int sc, sh, sw;
};

__device__ __forceinline__ Img32 cta_image(const ImgView &v, int b)
{
    return Img32{v.p + (long long)b * v.sb, (int)v.sc, (int)v.sh, (int)v.sw};
}

// ------------------------------------------------------------------------------------------------
// Division helpers (see header comment).
// ------------------------------------------------------------------------------------------------
template <bool IEEE>
__device__ __forceinline__ float div_const(float x, float d, float rcp)
{
    if (IEEE) return __fdiv_rn(x, d);
    const float q = __fmul_rn(x, rcp);
    const float r = __fmaf_rn(-d, q, x);
    return __fmaf_rn(r, rcp, q);
}

// u / d for a pixel coordinate u (d = W-1 or H-1 > 0).  |u| below 2^-122 needs no care: the caller
// computes (q - 0.5) * 2, which is -1 for any such q.  +-inf (z' == 0) must stay +-inf.
__device__ __forceinline__ float div_coord(float u, float d, float rcp, int exact)
{
    if (!exact) return __fdiv_rn(u, d);           // uniform branch: divisor failed the mantissa check
    const float q = __fmul_rn(u, rcp);
    const float r = __fmaf_rn(-d, q, u);
    const float q2 = __fmaf_rn(r, rcp, q);
    return (fabsf(u) == INFINITY) ? q : q2;
}

__device__ __forceinline__ bool value_out_of_fast_range(float v)
{
    const float a = fabsf(v);
    return !((a >= 0x1p-40f && a <= 0x1p40f) || a == 0.0f);    // NaN -> true
}

// ------------------------------------------------------------------------------------------------
// Camera constants of one batch element, staged once per CTA: cam[0..8] = inv_K[:3,:3],
// cam[9..20] = P = (K @ T)[:3, :] with the k-loop accumulated in order (unfused), like at::bmm's
// small-matrix path (view_synthesis.py:57).
// ------------------------------------------------------------------------------------------------
// bk = index of the intrinsics, bt = index of the pose (they differ when several source frames share a target)
__device__ __forceinline__ void stage_camera(const WPParams &p, int bk, int bt, float *cam)
{
    const int t = threadIdx.x;
    if (t < 9) {
        cam[t] = p.inv_K[bk * 16 + (t / 3) * 4 + (t % 3)];
    } else if (t < 21) {
        const int e = t - 9, i = e >> 2, j = e & 3;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; k++) acc = xadd(acc, xmul(p.K[bk * 16 + i * 4 + k], p.T[bt * 16 + k * 4 + j]));
        cam[t] = acc;
    }
}

struct Proj {
    float r0, r1, r2;      // inv_K[:3,:3] @ [x, y, 1]
    float X0, X1, X2;      // camera point
    float c0, c1, c2, z;   // P @ [X;1], z = c2 + eps
    float gx, gy, valid;   // normalised grid coordinate, validity
};

struct PixConst {          // per-thread copies of the hot scalars (keeps them out of the constant bank)
    float eps, wm1, hm1, half_w, half_h, rcpW, rcpH;
    int border, exact;
};

__device__ __forceinline__ PixConst pix_const(const WPParams &p)
{
    return PixConst{p.eps, p.wm1, p.hm1, p.half_w, p.half_h, p.rcpW, p.rcpH, p.border, p.div_exact};
}

// Pixel -> camera point -> projected homogeneous coordinate (everything before the divisions).
__device__ __forceinline__ void project_point(const float *cam, float eps, int x, int y, float d, Proj &o)
{
    const float fx = (float)x, fy = (float)y;
    // sgemm k-loop (k = 0,1,2) then * depth               view_synthesis.py:36-38
    o.r0 = xadd(xfma(cam[1], fy, xmul(cam[0], fx)), cam[2]);
    o.r1 = xadd(xfma(cam[4], fy, xmul(cam[3], fx)), cam[5]);
    o.r2 = xadd(xfma(cam[7], fy, xmul(cam[6], fx)), cam[8]);
    o.X0 = xmul(d, o.r0);
    o.X1 = xmul(d, o.r1);
    o.X2 = xmul(d, o.r2);
    const float *P = cam + 9;                              // view_synthesis.py:59
    o.c0 = xadd(xfma(P[2], o.X2, xfma(P[1], o.X1, xmul(P[0], o.X0))), P[3]);
    o.c1 = xadd(xfma(P[6], o.X2, xfma(P[5], o.X1, xmul(P[4], o.X0))), P[7]);
    o.c2 = xadd(xfma(P[10], o.X2, xfma(P[9], o.X1, xmul(P[8], o.X0))), P[11]);
    o.z = xadd(o.c2, eps);                                 // :60
}

__device__ __forceinline__ void project_pixel(const float *cam, const PixConst &k, int x, int y, float d, Proj &o)
{
    project_point(cam, k.eps, x, y, d, o);
    const float u = xdiv(o.c0, o.z), v = xdiv(o.c1, o.z);
    o.gx = xmul(xsub(div_coord(u, k.wm1, k.rcpW, k.exact), 0.5f), 2.0f);   // :66-68
    o.gy = xmul(xsub(div_coord(v, k.hm1, k.rcpH, k.exact), 0.5f), 2.0f);
    o.valid = (fabsf(o.gx) <= 1.0f && fabsf(o.gy) <= 1.0f) ? 1.0f : 0.0f;   // :70-71 (NaN -> 0)
}

// Bilinear sampling set-up, align_corners=False (ATen GridSamplerKernel.cpp, vectorised CPU path).
struct Samp {
    float ix, iy;              // after padding handling
    float nw, ne, sw, se;      // weights of taps (y0,x0) (y0,x1) (y1,x0) (y1,x1)
    int x0, y0;
    bool in00, in01, in10, in11;   // tap (row, col) in bounds: in<row><col>
    float mx, my;              // d(clamped)/d(unclamped): 0 where the border clamp is active
};

__device__ __forceinline__ void sampler_setup(const PixConst &k, float gx, float gy, Samp &s)
{
    float ix = xfma(xadd(gx, 1.0f), k.half_w, -0.5f);
    float iy = xfma(xadd(gy, 1.0f), k.half_h, -0.5f);
    s.mx = 1.0f;
    s.my = 1.0f;
    if (k.border) {
        s.mx = (ix > 0.0f && ix < k.wm1) ? 1.0f : 0.0f;    // clip_coordinates_set_grad
        s.my = (iy > 0.0f && iy < k.hm1) ? 1.0f : 0.0f;
        ix = fminf(k.wm1, fmaxf(0.0f, ix));                // NaN clamps to 0
        iy = fminf(k.hm1, fmaxf(0.0f, iy));
    }
    const float xw = floorf(ix), yn = floorf(iy);
    const float w = xsub(ix, xw), e = xsub(1.0f, w), n = xsub(iy, yn), so = xsub(1.0f, n);
    s.nw = xmul(so, e);
    s.ne = xmul(so, w);
    s.sw = xmul(n, e);
    s.se = xmul(n, w);
    s.ix = ix;
    s.iy = iy;
    // float comparisons so that NaN / huge coordinates are simply out of bounds
    const bool inx0 = (xw >= 0.0f) && (xw <= k.wm1), inx1 = (xw >= -1.0f) && (xw <= k.wm1 - 1.0f);
    const bool iny0 = (yn >= 0.0f) && (yn <= k.hm1), iny1 = (yn >= -1.0f) && (yn <= k.hm1 - 1.0f);
    s.in00 = iny0 && inx0;
    s.in01 = iny0 && inx1;
    s.in10 = iny1 && inx0;
    s.in11 = iny1 && inx1;
    s.x0 = (inx0 || inx1) ? (int)xw : 0;
    s.y0 = (iny0 || iny1) ? (int)yn : 0;
}

// Element offsets of the four taps (channel 0) in a 32-bit-strided image.
__device__ __forceinline__ void tap_offsets(const Img32 &im, const Samp &s, int o[4])
{
    o[0] = s.y0 * im.sh + s.x0 * im.sw;
    o[1] = o[0] + im.sw;
    o[2] = o[0] + im.sh;
    o[3] = o[2] + im.sw;
}

template <bool IL>
__device__ __forceinline__ void gather_taps(const Img32 &im, const Samp &s, const int o[4], int ch, float v[4])
{
    const int co = IL ? ch : ch * im.sc;
    v[0] = s.in00 ? __ldg(im.p + o[0] + co) : 0.0f;
    v[1] = s.in01 ? __ldg(im.p + o[1] + co) : 0.0f;
    v[2] = s.in10 ? __ldg(im.p + o[2] + co) : 0.0f;
    v[3] = s.in11 ? __ldg(im.p + o[3] + co) : 0.0f;
}

__device__ __forceinline__ float interp(const float v[4], const Samp &s)
{
    return xfma(v[3], s.se, xfma(v[2], s.sw, xfma(v[1], s.ne, xmul(v[0], s.nw))));
}

// ------------------------------------------------------------------------------------------------
// Shared-memory tile: NPL planes of RH x RP floats covering image rows [oy, oy+RH), cols [ox, ox+RP).
// reflect_fixup() fills the one-pixel ring just outside the image (row -1 <- row 1, row H <- row H-2,
// same for columns; corners via columns-then-rows), i.e. nn.ReflectionPad2d(1) (losses.py:18).
// The caller has synchronised after filling; the caller synchronises again afterwards.
// ------------------------------------------------------------------------------------------------
template <int NPL, int RH, int RP>
__device__ __forceinline__ void reflect_fixup(float *pl, int oy, int ox, int H, int W)
{
    const bool touches = (oy < 0) || (ox < 0) || (oy + RH > H) || (ox + RP > W);
    if (!touches) return;   // uniform per CTA
    for (int i = threadIdx.x; i < NPL * RH * 2; i += blockDim.x) {      // columns -1 and W, in-image rows
        const int side = i & 1, r = (i >> 1) % RH, k = (i >> 1) / RH;
        const int y = oy + r;
        if (y < 0 || y >= H) continue;
        const int lc = side ? (W - ox) : (-1 - ox);
        const int ls = side ? lc - 2 : lc + 2;
        if (lc < 0 || lc >= RP || ls < 0 || ls >= RP) continue;
        pl[(k * RH + r) * RP + lc] = pl[(k * RH + r) * RP + ls];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NPL * RP * 2; i += blockDim.x) {      // rows -1 and H, all columns
        const int side = i & 1, c = (i >> 1) % RP, k = (i >> 1) / RP;
        const int x = ox + c;
        if (x < -1 || x > W) continue;
        const int lr = side ? (H - oy) : (-1 - oy);
        const int ls = side ? lr - 2 : lr + 2;
        if (lr < 0 || lr >= RH || ls < 0 || ls >= RH) continue;
        pl[(k * RH + lr) * RP + c] = pl[(k * RH + ls) * RP + c];
    }
}

// ------------------------------------------------------------------------------------------------
// Exact-order 3x3 window sums for NP vertically adjacent centres (one channel).
// `wx`, `wy` point at the top-left sample of the first centre's window; S[p] = {Sx, Sy, Sxx, Syy, Sxy}
// accumulated in avg_pool2d's order (kh outer, kw inner, running fp32 sum).
// ------------------------------------------------------------------------------------------------
template <int NP, int RP>
__device__ __forceinline__ void window_sums(const float *wx, const float *wy, float (&S)[NP][5])
{
#pragma unroll
    for (int r = 0; r < NP + 2; r++) {
        float a[3], b[3], aa[3], bb[3], ab[3];
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
            a[dx] = wx[r * RP + dx];
            b[dx] = wy[r * RP + dx];
            aa[dx] = xmul(a[dx], a[dx]);
            bb[dx] = xmul(b[dx], b[dx]);
            ab[dx] = xmul(a[dx], b[dx]);
        }
#pragma unroll
        for (int pp = 0; pp < NP; pp++) {
            const int k = r - pp;   // row of centre pp's window
            if (k < 0 || k > 2) continue;
#pragma unroll
            for (int dx = 0; dx < 3; dx++) {
                if (k == 0 && dx == 0) {
                    S[pp][0] = a[0]; S[pp][1] = b[0]; S[pp][2] = aa[0]; S[pp][3] = bb[0]; S[pp][4] = ab[0];
                } else {
                    S[pp][0] = xadd(S[pp][0], a[dx]);
                    S[pp][1] = xadd(S[pp][1], b[dx]);
                    S[pp][2] = xadd(S[pp][2], aa[dx]);
                    S[pp][3] = xadd(S[pp][3], bb[dx]);
                    S[pp][4] = xadd(S[pp][4], ab[dx]);
                }
            }
        }
    }
}

struct SsimVals {
    float mux, muy, A1, A2, B1, B2, n, dn, Q, sraw, s, rdn;
};

template <bool IEEE>
__device__ __forceinline__ void ssim_finish(const float S[5], SsimVals &o)
{
    const float r9 = 1.0f / 9.0f;
    const float mux = div_const<IEEE>(S[0], 9.0f, r9), muy = div_const<IEEE>(S[1], 9.0f, r9);      // losses.py:27-28
    const float mxx = xmul(mux, mux), myy = xmul(muy, muy), mxy = xmul(mux, muy);
    const float vx = xsub(div_const<IEEE>(S[2], 9.0f, r9), mxx);                                     // :30-32
    const float vy = xsub(div_const<IEEE>(S[3], 9.0f, r9), myy);
    const float vxy = xsub(div_const<IEEE>(S[4], 9.0f, r9), mxy);
    o.A1 = xadd(xmul(xmul(2.0f, mux), muy), C1F);                               // :34
    o.A2 = xadd(xmul(2.0f, vxy), C2F);
    o.B1 = xadd(xadd(mxx, myy), C1F);                                           // :35
    o.B2 = xadd(xadd(vx, vy), C2F);
    o.n = xmul(o.A1, o.A2);
    o.dn = xmul(o.B1, o.B2);
    o.Q = xdiv(o.n, o.dn);
    o.sraw = xmul(xsub(1.0f, o.Q), 0.5f);                                       // :37  (/2 is exact)
    o.s = o.sraw < 0.0f ? 0.0f : (o.sraw > 1.0f ? 1.0f : o.sraw);               // NaN propagates
    o.mux = mux;
    o.muy = muy;
}

// ------------------------------------------------------------------------------------------------
// Phase 1: fill x / y planes for the region.  WARP: x = syn (*valid), y = tgt (*valid).
// Returns true if this thread stored a value outside the fast-division range.
// ------------------------------------------------------------------------------------------------
template <int MODE, int CK, int RH, int RP, int HALO, int TH, int TW, bool IL>
__device__ __forceinline__ bool fill_region(const WPParams &p, const float *cam, int b, int ch0, int ty0, int tx0,
                                            float *sx, float *sy, bool write_outputs)
{
    const int oy = ty0 - HALO, ox = tx0 - HALO;
    const int H = p.H, W = p.W;
    const Img32 src = cta_image(p.src, b), tgt = cta_image(p.tgt, b);
    const PixConst k = pix_const(p);
    const bool use_mask = p.use_mask != 0;
    const float *depth_b = p.depth + (long long)b * H * W;
    bool bad = false;
    for (int i = threadIdx.x; i < RH * RP; i += blockDim.x) {
        const int hy = i / RP, hx = i - hy * RP;
        const int y = oy + hy, x = ox + hx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        float *px = sx + hy * RP + hx, *py = sy + hy * RP + hx;
        if (MODE == MODE_WARP) {
            const int pixo = y * W + x;
            const float d = __ldg(depth_b + pixo);
            Proj pr;
            project_pixel(cam, k, x, y, d, pr);
            Samp s;
            sampler_setup(k, pr.gx, pr.gy, s);
            int o[4];
            tap_offsets(src, s, o);
            const float *tp = tgt.p + y * tgt.sh + x * tgt.sw;
            const bool centre = write_outputs && hy >= HALO && hy < HALO + TH && hx >= HALO && hx < HALO + TW;
            if (centre) {
                const long long pixi = (long long)b * H * W + pixo;
                if (p.valid) p.valid[pixi] = pr.valid;
                if (p.pix) { p.pix[pixi * 2] = pr.gx; p.pix[pixi * 2 + 1] = pr.gy; }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                float v[4];
                gather_taps<IL>(src, s, o, ch, v);
                const float sv = interp(v, s);
                const float t = __ldg(tp + (IL ? ch : ch * tgt.sc));
                const float xv = use_mask ? xmul(sv, pr.valid) : sv;        // train_depth.py:714-715
                const float yv = use_mask ? xmul(t, pr.valid) : t;
                px[ch * RH * RP] = xv;
                py[ch * RH * RP] = yv;
                bad |= value_out_of_fast_range(xv) | value_out_of_fast_range(yv);
                if (centre && p.syn) p.syn[((long long)b * 3 + ch) * H * W + pixo] = sv;
            }
        } else {
            const float *xp = src.p + y * src.sh + x * src.sw, *yp = tgt.p + y * tgt.sh + x * tgt.sw;
#pragma unroll
            for (int ch = 0; ch < CK; ch++) {
                const float xv = __ldg(xp + (long long)(ch0 + ch) * src.sc), yv = __ldg(yp + (long long)(ch0 + ch) * tgt.sc);
                px[ch * RH * RP] = xv;
                py[ch * RH * RP] = yv;
                bad |= value_out_of_fast_range(xv) | value_out_of_fast_range(yv);
            }
        }
    }
    return bad;
}

// Host helpers defined in warp_photo.cu
bool view_fits_int32(const ImgView &v, int C, int H, int W);
int fill_common(WPParams &p, int B, int C, int H, int W, int padding_mode, int use_mask, float eps, cudaStream_t st);
int set_views(WPParams &p, const float *a, const int64_t as[4], const float *b, const int64_t bs[4], int C);
int launch_reduce_partials(const float *partial, long long n, double scale, float *out, cudaStream_t st);
int launch_reduce_gP(const float *partial, int ctas_per_b, int B, float *gP, cudaStream_t st, const float *skip_flag = nullptr);
int launch_reduce_loss_gP(const float *loss_partial, long long n, double scale, float *loss, const float *gp_partial, int ctas_per_b, int B,
                          float *gP, cudaStream_t st, const float *skip_flag = nullptr);
// Defined in warp_photo_fused.cu: the streaming value + gradient kernel on a prepared parameter block
size_t stream_workspace_bytes(int B, int H, int W);
int launch_stream(WPParams &p, int B, int H, int W, float *loss_mean, float *grad_P, void *workspace, size_t workspace_bytes, cudaStream_t st);
int launch_ssim_stream_bwd(WPParams &p, int B, int H, int W, cudaStream_t st);


}  // namespace e2e
