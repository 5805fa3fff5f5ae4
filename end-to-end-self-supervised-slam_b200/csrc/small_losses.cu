// Small fused losses of loss/losses.py: edge-aware disparity smoothness, sparse ground-truth L1,
// depth regulariser, geometric consistency.  Each is a grid-stride pass (grid = a multiple of the SM
// count) producing per-CTA partial sums in double, followed by a one-CTA deterministic final reduce;
// nothing synchronises with the host.
#include <cstdint>

#include "common.cuh"

namespace e2e {

constexpr int RED_NT = 256;
constexpr int RED_MAX_BLOCKS = kNumSMs * 4;

static inline int red_blocks(long long n)
{
    long long b = (n + RED_NT - 1) / RED_NT;
    if (b < 1) b = 1;
    return (int)(b > RED_MAX_BLOCKS ? RED_MAX_BLOCKS : b);
}

// block-level sum of NV doubles per thread -> partial[blockIdx.x * NV + k]
template <int NV>
__device__ __forceinline__ void block_partials(double (&v)[NV], double *partial)
{
    __shared__ double sh[NV][RED_NT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        v[k] = warp_sum_d(v[k]);
        if (lane == 0) sh[k][wid] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int w = 0; w < RED_NT / 32; w++) t += sh[threadIdx.x][w];
        partial[(long long)blockIdx.x * NV + threadIdx.x] = t;
    }
}

// sum of partial[i*NV + k] over i < n, by one warp-sized team of the calling block; result broadcast.
template <int NV>
__device__ __forceinline__ void sum_partials(const double *partial, int n, double (&out)[NV])
{
    __shared__ double tot[NV];
    __syncthreads();                       // a previous call's readers are done with tot[]
    if (threadIdx.x < 32) {
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) acc[k] = 0.0;
        for (int i = threadIdx.x; i < n; i += 32)
#pragma unroll
            for (int k = 0; k < NV; k++) acc[k] += partial[(long long)i * NV + k];
#pragma unroll
        for (int k = 0; k < NV; k++) {
            acc[k] = warp_sum_d(acc[k]);
            if (threadIdx.x == 0) tot[k] = acc[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) out[k] = tot[k];
}

// ------------------------------------------------------------------------------------------------
// Smoothness (train_depth.py:763-773 + losses.py:119-132)
// ------------------------------------------------------------------------------------------------
// pass 1: per-image sum of disp.  grid (blocks, B); partial[b][blk]
__global__ void __launch_bounds__(RED_NT) smooth_mean_kernel(const float *disp, int HW, double *partial)
{
    const int b = blockIdx.y;
    const float *d = disp + (long long)b * HW;
    double v[1] = {0.0};
    for (int i = blockIdx.x * RED_NT + threadIdx.x; i < HW; i += gridDim.x * RED_NT) v[0] += (double)d[i];
    block_partials<1>(v, partial + (long long)b * gridDim.x);
}

struct SmoothParams {
    const float *disp;
    ImgView img;
    int B, H, W, nblk_mean;
    const double *mean_partial;     // [B][nblk_mean]
    double *partial;                // forward: [blocks*B][2]; backward pass 1: [B][blocks][1]
    const float *grad_loss;
    float *grad_disp;
    const double *dot_partial;
    int nblk_dot;
};

__device__ __forceinline__ float img_edge(const ImgView &im, long long o0, long long o1)
{
    // mean_c |I0 - I1|, channels summed in order then /3 (losses.py:125-126)
    const float a = fabsf(im.p[o0] - im.p[o1]);
    const float b = fabsf(im.p[o0 + im.sc] - im.p[o1 + im.sc]);
    const float c = fabsf(im.p[o0 + 2 * im.sc] - im.p[o1 + 2 * im.sc]);
    return ((a + b) + c) / 3.0f;
}

// pass 2 (forward): sum of x-terms and y-terms.  grid (blocks, B)
__global__ void __launch_bounds__(RED_NT) smooth_fwd_kernel(const SmoothParams p)
{
    const int b = blockIdx.y, H = p.H, W = p.W, HW = H * W;
    double ms[1];
    sum_partials<1>(p.mean_partial + (long long)b * p.nblk_mean, p.nblk_mean, ms);
    const float m = (float)(ms[0] / (double)HW);
    const float me = xadd(m, 1e-7f);                      // norm = disp / (mean + 1e-7), train_depth.py:769
    const float *d = p.disp + (long long)b * HW;
    const long long ib = (long long)b * p.img.sb;
    double v[2] = {0.0, 0.0};
    for (int i = blockIdx.x * RED_NT + threadIdx.x; i < HW; i += gridDim.x * RED_NT) {
        const int y = i / W, x = i - y * W;
        const float n0 = xdiv(d[i], me);
        const long long o0 = ib + (long long)y * p.img.sh + (long long)x * p.img.sw;
        if (x + 1 < W) v[0] += (double)(fabsf(xsub(n0, xdiv(d[i + 1], me))) * expf(-img_edge(p.img, o0, o0 + p.img.sw)));
        if (y + 1 < H) v[1] += (double)(fabsf(xsub(n0, xdiv(d[i + W], me))) * expf(-img_edge(p.img, o0, o0 + p.img.sh)));
    }
    block_partials<2>(v, p.partial + ((long long)b * gridDim.x) * 2);
}

__global__ void __launch_bounds__(RED_NT) smooth_final_kernel(const double *partial, int n, double inv_nx, double inv_ny, float *loss)
{
    double t[2];
    sum_partials<2>(partial, n, t);
    if (threadIdx.x == 0) loss[0] = (float)(t[0] * inv_nx + t[1] * inv_ny);
}

// dL/dn at pixel i (L = mean_x-terms + mean_y-terms), gather form over the four pairs touching i.
// Every normalised value is the explicitly rounded quotient disp / (mean + 1e-7) and every difference an
// explicitly rounded subtraction, so the two pixels of a pair always agree on sign(n_i - n_j) (a fused
// multiply-subtract would not: neighbours one ulp apart collapse to the same quotient).
__device__ __forceinline__ float smooth_gn(const SmoothParams &p, const float *d, long long ib, int y, int x, float me,
                                           float cx, float cy)
{
    const int W = p.W, H = p.H, i = y * W + x;
    const float n0 = xdiv(d[i], me);
    const long long o0 = ib + (long long)y * p.img.sh + (long long)x * p.img.sw;
    float g = 0.f;
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    if (x + 1 < W) g += cx * sgn(xsub(n0, xdiv(d[i + 1], me))) * expf(-img_edge(p.img, o0, o0 + p.img.sw));
    if (x > 0) g -= cx * sgn(xsub(xdiv(d[i - 1], me), n0)) * expf(-img_edge(p.img, o0 - p.img.sw, o0));
    if (y + 1 < H) g += cy * sgn(xsub(n0, xdiv(d[i + W], me))) * expf(-img_edge(p.img, o0, o0 + p.img.sh));
    if (y > 0) g -= cy * sgn(xsub(xdiv(d[i - W], me), n0)) * expf(-img_edge(p.img, o0 - p.img.sh, o0));
    return g;
}

// backward pass 1: per-image dot = sum_j gn_j * disp_j.   pass 2: grad = gn/(m+e) - dot/((m+e)^2 HW)
template <int PASS>
__global__ void __launch_bounds__(RED_NT) smooth_bwd_kernel(const SmoothParams p)
{
    const int b = blockIdx.y, H = p.H, W = p.W, HW = H * W;
    double ms[1];
    sum_partials<1>(p.mean_partial + (long long)b * p.nblk_mean, p.nblk_mean, ms);
    const float m = (float)(ms[0] / (double)HW);
    const float me = xadd(m, 1e-7f), inv = 1.0f / me;
    const float up = p.grad_loss ? p.grad_loss[0] : 1.0f;
    const float cx = up / (float)((double)p.B * H * (W - 1)), cy = up / (float)((double)p.B * (H - 1) * W);
    const float *d = p.disp + (long long)b * HW;
    const long long ib = (long long)b * p.img.sb;
    if (PASS == 1) {
        double v[1] = {0.0};
        for (int i = blockIdx.x * RED_NT + threadIdx.x; i < HW; i += gridDim.x * RED_NT) {
            const int y = i / W, x = i - y * W;
            v[0] += (double)smooth_gn(p, d, ib, y, x, me, cx, cy) * (double)d[i];
        }
        block_partials<1>(v, p.partial + (long long)b * gridDim.x);
    } else {
        double ds[1];
        sum_partials<1>(p.dot_partial + (long long)b * p.nblk_dot, p.nblk_dot, ds);
        const float corr = (float)(ds[0] / ((double)me * (double)me * (double)HW));
        for (int i = blockIdx.x * RED_NT + threadIdx.x; i < HW; i += gridDim.x * RED_NT) {
            const int y = i / W, x = i - y * W;
            p.grad_disp[(long long)b * HW + i] = smooth_gn(p, d, ib, y, x, me, cx, cy) * inv - corr;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Elementwise reductions: sparse-gt L1 (losses.py:151-160), regulariser (:134-148), geometric (:84-95)
// ------------------------------------------------------------------------------------------------
enum { EW_SPARSE_L1 = 0, EW_REG_L1 = 1, EW_REG_L2 = 2, EW_GEOMETRIC = 3 };

template <int KIND>
__global__ void __launch_bounds__(RED_NT) ew_fwd_kernel(const float *a, const float *b, const float *c, long long n, double *partial)
{
    double v[2] = {0.0, 0.0};
    auto term = [&](float ai, float bi, float ci) {
        if (KIND == EW_SPARSE_L1) v[0] += (double)fabsf(ai * bi - ci);                      // a=pred b=mask c=gt
        if (KIND == EW_REG_L1) v[0] += (double)fabsf(ai - bi);                              // a=initial b=refined
        if (KIND == EW_REG_L2) { const float t = ai - bi; v[0] += (double)(t * t); }
        if (KIND == EW_GEOMETRIC) {                                                          // a=warped b=interp c=valid
            float t = fabsf(ai - bi) / (ai + bi);
            t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);                                        // NaN stays NaN like torch.clamp
            v[0] += (double)(t * ci);
            v[1] += (double)ci;
        }
    };
    const bool has_c = (KIND == EW_SPARSE_L1 || KIND == EW_GEOMETRIC);
    if ((n & 3) == 0 && ((((uintptr_t)a) | ((uintptr_t)b) | (has_c ? (uintptr_t)c : 0)) & 15u) == 0) {      // 16-byte loads
        const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b), *c4 = reinterpret_cast<const float4 *>(c);
        for (long long i = (long long)blockIdx.x * RED_NT + threadIdx.x; i < n / 4; i += (long long)gridDim.x * RED_NT) {
            const float4 x = a4[i], y = b4[i], z = has_c ? c4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            term(x.x, y.x, z.x); term(x.y, y.y, z.y); term(x.z, y.z, z.z); term(x.w, y.w, z.w);
        }
    } else {
        for (long long i = (long long)blockIdx.x * RED_NT + threadIdx.x; i < n; i += (long long)gridDim.x * RED_NT)
            term(a[i], b[i], has_c ? c[i] : 0.f);
    }
    block_partials<2>(v, partial);
}

template <int KIND>
__global__ void __launch_bounds__(RED_NT) ew_final_kernel(const double *partial, int nblk, double inv_n, float *loss)
{
    double t[2];
    sum_partials<2>(partial, nblk, t);
    if (threadIdx.x == 0) {
        if (KIND == EW_GEOMETRIC) loss[0] = (t[1] > 10000.0) ? (float)(t[0] / t[1]) : 0.0f;   // losses.py:90
        else loss[0] = (float)(t[0] * inv_n);
    }
}

template <int KIND>
__global__ void __launch_bounds__(RED_NT) ew_bwd_kernel(const float *a, const float *b, const float *c, long long n,
                                                        const float *grad_loss, float *grad)
{
    const float g = (grad_loss ? grad_loss[0] : 1.0f) / (float)n;
    auto term = [&](float ai, float bi, float ci) -> float {
        if (KIND == EW_SPARSE_L1) {
            const float t = ai * bi - ci;
            return g * (t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f)) * bi;
        }
        if (KIND == EW_REG_L1) {
            const float t = bi - ai;
            return g * (t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f));
        }
        return g * 2.0f * (bi - ai);      // EW_REG_L2
    };
    const bool has_c = (KIND == EW_SPARSE_L1);
    if ((n & 3) == 0 && ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)grad) | (has_c ? (uintptr_t)c : 0)) & 15u) == 0) {      // 16-byte accesses
        const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b), *c4 = reinterpret_cast<const float4 *>(c);
        float4 *g4 = reinterpret_cast<float4 *>(grad);
        for (long long i = (long long)blockIdx.x * RED_NT + threadIdx.x; i < n / 4; i += (long long)gridDim.x * RED_NT) {
            const float4 x = a4[i], y = b4[i], z = has_c ? c4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            g4[i] = make_float4(term(x.x, y.x, z.x), term(x.y, y.y, z.y), term(x.z, y.z, z.z), term(x.w, y.w, z.w));
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * RED_NT + threadIdx.x; i < n; i += (long long)gridDim.x * RED_NT)
        grad[i] = term(a[i], b[i], has_c ? c[i] : 0.f);
}

// d loss / d {warped, interpolated} of geometric_consistency_loss: loss = sum(t * m) / sum(m) if sum(m) > 10000 else 0,
// t = clamp(|a - b| / (a + b), 0, 1) (clamp passes the gradient on the closed interval, like torch.clamp)
__global__ void __launch_bounds__(RED_NT) geometric_bwd_kernel(const float *a, const float *b, const float *m, long long n,
                                                               const float *mask_sum, const float *grad_loss, float *ga, float *gb)
{
    const float cnt = mask_sum[0];
    const float g = (cnt > 10000.0f) ? (grad_loss ? grad_loss[0] : 1.0f) / cnt : 0.0f;
    for (long long i = (long long)blockIdx.x * RED_NT + threadIdx.x; i < n; i += (long long)gridDim.x * RED_NT) {
        const float d = a[i] - b[i], s = a[i] + b[i], t = fabsf(d) / s;
        float da = 0.f, db = 0.f;
        if (g != 0.0f && t >= 0.f && t <= 1.f) {
            const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f), q = fabsf(d) / (s * s);
            da = g * m[i] * (sg / s - q);
            db = g * m[i] * (-sg / s - q);
        }
        if (ga) ga[i] = da;
        if (gb) gb[i] = db;
    }
}

static int check_ws(void *ws, size_t bytes, size_t need)
{
    E2E_REQUIRE(ws && bytes >= need, "workspace too small: need %zu bytes, got %zu", need, bytes);
    E2E_REQUIRE(((uintptr_t)ws & 7) == 0, "workspace must be 8-byte aligned");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Per-pixel minimum over up to 8 candidate loss maps + mean (train_depth.py:653-658: `torch.min(photmetric, dim=1)` then
// `.mean()`): the min-reprojection / auto-masking objective.  Candidates are separate [B,1,H,W] maps (the per-source photometric
// maps and the identity maps), so the reference's torch.cat is never materialised.  torch.min semantics: the FIRST minimal
// candidate wins, a NaN candidate wins over everything after it is met.  The winning index is kept (uint8) for the backward pass,
// which routes the upstream gradient to that candidate only.
// ------------------------------------------------------------------------------------------------
struct MinCands {
    const float *p[8];
    float *g[8];
    int C;
};

__global__ void __launch_bounds__(RED_NT) min_composite_fwd_kernel(const MinCands c, long long n, unsigned char *index, double *partial)
{
    double acc[1] = {0.0};
    for (long long i = (long long)blockIdx.x * RED_NT + threadIdx.x; i < n; i += (long long)gridDim.x * RED_NT) {
        float best = __ldg(c.p[0] + i);
        int bi = 0;
#pragma unroll
        for (int k = 1; k < 8; k++) {
            if (k < c.C) {
                const float v = __ldg(c.p[k] + i);
                if (!(best != best) && (v < best || v != v)) { best = v; bi = k; }
            }
        }
        index[i] = (unsigned char)bi;
        acc[0] += (double)best;
    }
    block_partials<1>(acc, partial);
}

__global__ void __launch_bounds__(RED_NT) min_composite_final_kernel(const double *partial, int nblk, double inv_n, float *loss)
{
    double t[1];
    sum_partials<1>(partial, nblk, t);
    if (threadIdx.x == 0) loss[0] = (float)(t[0] * inv_n);
}

__global__ void __launch_bounds__(RED_NT) min_composite_bwd_kernel(const MinCands c, long long n, const unsigned char *index,
                                                                   const float *grad_loss, float scale)
{
    const float gs = scale * (grad_loss ? __ldg(grad_loss) : 1.0f);
    for (long long i = (long long)blockIdx.x * RED_NT + threadIdx.x; i < n; i += (long long)gridDim.x * RED_NT) {
        const int bi = index[i];
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < c.C && c.g[k]) c.g[k][i] = (k == bi) ? gs : 0.0f;
    }
}

}  // namespace e2e

using namespace e2e;

static int smooth_setup(SmoothParams &p, const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                        void *workspace, size_t workspace_bytes, int &blocks)
{
    E2E_REQUIRE(disp && img && B > 0 && H >= 2 && W >= 2, "smoothness: bad arguments");
    blocks = red_blocks((long long)H * W);
    if (blocks * B > RED_MAX_BLOCKS) blocks = RED_MAX_BLOCKS / B > 0 ? RED_MAX_BLOCKS / B : 1;
    E2E_REQUIRE(B <= RED_MAX_BLOCKS, "smoothness: batch larger than %d", RED_MAX_BLOCKS);
    if (int rc = check_ws(workspace, workspace_bytes, sizeof(double) * (size_t)blocks * B * 4)) return rc;
    p.disp = disp; p.img = make_view(img, img_strides); p.B = B; p.H = H; p.W = W;
    p.nblk_mean = blocks;
    p.mean_partial = (const double *)workspace;
    return 0;
}

template <int KIND>
static int ew_fwd(const float *a, const float *b, const float *c, long long n, float *loss, void *workspace, size_t workspace_bytes,
                  cudaStream_t st)
{
    E2E_REQUIRE(a && b && loss && n > 0, "elementwise loss: bad arguments");
    const int blocks = red_blocks(n);
    if (int rc = check_ws(workspace, workspace_bytes, sizeof(double) * (size_t)blocks * 2)) return rc;
    ew_fwd_kernel<KIND><<<blocks, RED_NT, 0, st>>>(a, b, c, n, (double *)workspace);
    ew_final_kernel<KIND><<<1, RED_NT, 0, st>>>((const double *)workspace, blocks, 1.0 / (double)n, loss);
    count_launch(2);
    return finish_launch("elementwise loss");
}


extern "C" {

size_t e2e_reduce_workspace_bytes(long long n_elements)
{
    (void)n_elements;
    // mean partials + main partials (2 doubles per CTA) with batch folded into the CTA count, generous bound
    return sizeof(double) * (size_t)RED_MAX_BLOCKS * 8 + 256;
}

int e2e_smooth_fwd(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                   float *loss, void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    SmoothParams p = {};
    int blocks;
    if (int rc = smooth_setup(p, disp, img, img_strides, B, H, W, workspace, workspace_bytes, blocks)) return rc;
    double *ws = (double *)workspace;
    p.partial = ws + (size_t)blocks * B;
    smooth_mean_kernel<<<dim3(blocks, B), RED_NT, 0, st>>>(disp, H * W, ws);
    smooth_fwd_kernel<<<dim3(blocks, B), RED_NT, 0, st>>>(p);
    smooth_final_kernel<<<1, RED_NT, 0, st>>>(p.partial, blocks * B, 1.0 / ((double)B * H * (W - 1)), 1.0 / ((double)B * (H - 1) * W), loss);
    count_launch(3);
    return finish_launch("smooth_fwd");
}

int e2e_smooth_bwd(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                   const float *grad_loss, float *grad_disp, void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    SmoothParams p = {};
    int blocks;
    if (int rc = smooth_setup(p, disp, img, img_strides, B, H, W, workspace, workspace_bytes, blocks)) return rc;
    E2E_REQUIRE(grad_disp, "smoothness: null grad_disp");
    double *ws = (double *)workspace;
    p.partial = ws + (size_t)blocks * B;
    p.dot_partial = p.partial;
    p.nblk_dot = blocks;
    p.grad_loss = grad_loss;
    p.grad_disp = grad_disp;
    smooth_mean_kernel<<<dim3(blocks, B), RED_NT, 0, st>>>(disp, H * W, ws);
    smooth_bwd_kernel<1><<<dim3(blocks, B), RED_NT, 0, st>>>(p);
    smooth_bwd_kernel<2><<<dim3(blocks, B), RED_NT, 0, st>>>(p);
    count_launch(3);
    return finish_launch("smooth_bwd");
}

int e2e_sparse_l1_fwd(const float *pred, const float *mask, const float *gt, long long n, float *loss,
                      void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(gt, "sparse_l1: null gt");
    return ew_fwd<EW_SPARSE_L1>(pred, mask, gt, n, loss, workspace, workspace_bytes, (cudaStream_t)stream);
}

int e2e_sparse_l1_bwd(const float *pred, const float *mask, const float *gt, long long n,
                      const float *grad_loss, float *grad_pred, void *stream)
{
    E2E_REQUIRE(pred && mask && gt && grad_pred && n > 0, "sparse_l1_bwd: bad arguments");
    ew_bwd_kernel<EW_SPARSE_L1><<<red_blocks(n), RED_NT, 0, (cudaStream_t)stream>>>(pred, mask, gt, n, grad_loss, grad_pred);
    count_launch();
    return finish_launch("sparse_l1_bwd");
}

int e2e_depth_reg_fwd(const float *initial, const float *refined, long long n, int kind, float *loss,
                      void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(kind == 1 || kind == 2, "please specify a correct norm");     /* losses.py:146 */
    if (kind == 1) return ew_fwd<EW_REG_L1>(initial, refined, nullptr, n, loss, workspace, workspace_bytes, (cudaStream_t)stream);
    return ew_fwd<EW_REG_L2>(initial, refined, nullptr, n, loss, workspace, workspace_bytes, (cudaStream_t)stream);
}

int e2e_depth_reg_bwd(const float *initial, const float *refined, long long n, int kind,
                      const float *grad_loss, float *grad_refined, void *stream)
{
    E2E_REQUIRE(initial && refined && grad_refined && n > 0 && (kind == 1 || kind == 2), "depth_reg_bwd: bad arguments");
    if (kind == 1) ew_bwd_kernel<EW_REG_L1><<<red_blocks(n), RED_NT, 0, (cudaStream_t)stream>>>(initial, refined, nullptr, n, grad_loss, grad_refined);
    else ew_bwd_kernel<EW_REG_L2><<<red_blocks(n), RED_NT, 0, (cudaStream_t)stream>>>(initial, refined, nullptr, n, grad_loss, grad_refined);
    count_launch();
    return finish_launch("depth_reg_bwd");
}

int e2e_geometric_fwd(const float *warped_depth, const float *interp_depth, const float *valid, long long n,
                      float *loss, void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(valid, "geometric: null valid mask");
    return ew_fwd<EW_GEOMETRIC>(warped_depth, interp_depth, valid, n, loss, workspace, workspace_bytes, (cudaStream_t)stream);
}

int e2e_geometric_bwd(const float *warped_depth, const float *interp_depth, const float *valid, long long n, const float *mask_sum,
                      const float *grad_loss, float *grad_warped, float *grad_interp, void *stream)
{
    E2E_REQUIRE(warped_depth && interp_depth && valid && mask_sum && (grad_warped || grad_interp) && n > 0, "geometric_bwd: bad arguments");
    geometric_bwd_kernel<<<red_blocks(n), RED_NT, 0, (cudaStream_t)stream>>>(warped_depth, interp_depth, valid, n, mask_sum, grad_loss,
                                                                             grad_warped, grad_interp);
    count_launch();
    return finish_launch("geometric_bwd");
}

int e2e_min_composite_fwd(const float *const *candidates, int n_candidates, long long n, unsigned char *index, float *loss,
                          void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(candidates && n_candidates >= 1 && n_candidates <= 8 && n > 0 && index && loss, "min_composite: bad arguments");
    MinCands c = {};
    c.C = n_candidates;
    for (int k = 0; k < n_candidates; k++) {
        E2E_REQUIRE(candidates[k], "min_composite: null candidate map");
        c.p[k] = candidates[k];
    }
    const int blocks = red_blocks(n);
    E2E_REQUIRE(workspace && workspace_bytes >= blocks * sizeof(double), "min_composite: workspace too small");
    min_composite_fwd_kernel<<<blocks, RED_NT, 0, (cudaStream_t)stream>>>(c, n, index, (double *)workspace);
    min_composite_final_kernel<<<1, RED_NT, 0, (cudaStream_t)stream>>>((const double *)workspace, blocks, 1.0 / (double)n, loss);
    count_launch(2);
    return finish_launch("min_composite_fwd");
}

int e2e_min_composite_bwd(const unsigned char *index, int n_candidates, long long n, const float *grad_loss, float *const *grad_candidates,
                          void *stream)
{
    E2E_REQUIRE(index && grad_candidates && n_candidates >= 1 && n_candidates <= 8 && n > 0, "min_composite_bwd: bad arguments");
    MinCands c = {};
    c.C = n_candidates;
    for (int k = 0; k < n_candidates; k++) c.g[k] = grad_candidates[k];
    min_composite_bwd_kernel<<<red_blocks(n), RED_NT, 0, (cudaStream_t)stream>>>(c, n, index, grad_loss, (float)(1.0 / (double)n));
    count_launch();
    return finish_launch("min_composite_bwd");
}

}  // extern "C"
