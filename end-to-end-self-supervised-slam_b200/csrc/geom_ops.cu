// Granular view-synthesis ops (tier (i) of SURVEY.md section 8(b)): one kernel per reference call with
// the reference's own tensors in and out, for scripts that are not switched to the fused op.
//   BackprojectDepth.forward  depth_estimation/view_synthesis.py:34-40
//   Project3D.forward         depth_estimation/view_synthesis.py:54-78
//   F.grid_sample (bilinear; zeros|border; align_corners False|True)   train_depth.py:568-590
// Forward arithmetic is in the reference's operation order (see oracle/warp_photo_oracle.c).
#include "common.cuh"

namespace e2e {

constexpr int G_NT = 256;

static inline int ew_grid(long long n)
{
    long long b = (n + G_NT - 1) / G_NT;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- backproject ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(G_NT) backproject_fwd_kernel(const float *depth, const float *inv_K, int B, int H, int W, float *cam)
{
    const long long HW = (long long)H * W, n = (long long)B * HW;
    for (long long i = (long long)blockIdx.x * G_NT + threadIdx.x; i < n; i += (long long)gridDim.x * G_NT) {
        const int b = (int)(i / HW);
        const long long j = i - (long long)b * HW;
        const int y = (int)(j / W), x = (int)(j - (long long)y * W);
        const float *ik = inv_K + b * 16;
        const float fx = (float)x, fy = (float)y, d = depth[i];
        float *o = cam + (long long)b * 4 * HW + j;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const float ray = xadd(xfma(ik[r * 4 + 1], fy, xmul(ik[r * 4 + 0], fx)), ik[r * 4 + 2]);
            o[r * HW] = xmul(d, ray);
        }
        o[3 * HW] = 1.0f;
    }
}

__global__ void __launch_bounds__(G_NT) backproject_bwd_kernel(const float *gcam, const float *inv_K, int B, int H, int W, float *gdepth)
{
    const long long HW = (long long)H * W, n = (long long)B * HW;
    for (long long i = (long long)blockIdx.x * G_NT + threadIdx.x; i < n; i += (long long)gridDim.x * G_NT) {
        const int b = (int)(i / HW);
        const long long j = i - (long long)b * HW;
        const int y = (int)(j / W), x = (int)(j - (long long)y * W);
        const float *ik = inv_K + b * 16;
        const float fx = (float)x, fy = (float)y;
        const float *g = gcam + (long long)b * 4 * HW + j;
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const float ray = xadd(xfma(ik[r * 4 + 1], fy, xmul(ik[r * 4 + 0], fx)), ik[r * 4 + 2]);
            acc += g[r * HW] * ray;
        }
        gdepth[i] = acc;
    }
}

// ---- project3d -----------------------------------------------------------------------------------
struct P3Params {
    const float *points, *K, *T;
    int B, H, W, exact;
    float eps, wm1, hm1, rcpW, rcpH;
    float *pix, *valid, *wdepth;
    const float *g_pix, *g_wdepth;
    float *g_points, *gP_partial;
};

__device__ __forceinline__ void compose_P(const float *K, const float *T, float *P)   // thread < 12
{
    const int e = threadIdx.x, i = e >> 2, j = e & 3;
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) acc = xadd(acc, xmul(K[i * 4 + k], T[k * 4 + j]));
    P[e] = acc;
}

__device__ __forceinline__ float p3_div(float u, float d, float rcp, int exact)
{
    if (!exact) return __fdiv_rn(u, d);
    const float q = __fmul_rn(u, rcp);
    const float r = __fmaf_rn(-d, q, u);
    const float q2 = __fmaf_rn(r, rcp, q);
    // unlike the fused kernel, pix is an OUTPUT here, so the tiny range must be exact too
    const float a = fabsf(u);
    return (a == INFINITY) ? q : ((a < 0x1p-120f && a != 0.0f) ? __fdiv_rn(u, d) : q2);
}

template <bool BWD>
__global__ void __launch_bounds__(G_NT) project3d_kernel(const P3Params p)
{
    __shared__ float P[12];
    __shared__ float red[(G_NT / 32) * 12];
    const int b = blockIdx.y;
    const int HW = p.H * p.W;
    if (threadIdx.x < 12) compose_P(p.K + b * 16, p.T + b * 16, P);
    __syncthreads();
    const float *pts = p.points + (long long)b * 4 * HW;
    float gP[12];
#pragma unroll
    for (int e = 0; e < 12; e++) gP[e] = 0.f;
    for (int j = blockIdx.x * G_NT + threadIdx.x; j < HW; j += gridDim.x * G_NT) {
        const float X0 = pts[j], X1 = pts[HW + j], X2 = pts[2 * HW + j], X3 = pts[3 * HW + j];
        float c[3];
#pragma unroll
        for (int i = 0; i < 3; i++)      // sgemm k-loop over the four homogeneous coordinates
            c[i] = xfma(P[i * 4 + 3], X3, xfma(P[i * 4 + 2], X2, xfma(P[i * 4 + 1], X1, xmul(P[i * 4 + 0], X0))));
        const float z = xadd(c[2], p.eps);
        const float u = xdiv(c[0], z), v = xdiv(c[1], z);
        const long long o = (long long)b * HW + j;
        if (!BWD) {
            const float gx = xmul(xsub(p3_div(u, p.wm1, p.rcpW, p.exact), 0.5f), 2.0f);
            const float gy = xmul(xsub(p3_div(v, p.hm1, p.rcpH, p.exact), 0.5f), 2.0f);
            p.pix[o * 2] = gx;
            p.pix[o * 2 + 1] = gy;
            p.valid[o] = (fabsf(gx) <= 1.0f && fabsf(gy) <= 1.0f) ? 1.0f : 0.0f;
            if (p.wdepth) p.wdepth[o] = fmaxf(c[2], 1e-3f);                       // clamp(min=1e-3), view_synthesis.py:74
        } else {
            const float gu = p.g_pix ? p.g_pix[o * 2] * 2.0f / p.wm1 : 0.f, gv = p.g_pix ? p.g_pix[o * 2 + 1] * 2.0f / p.hm1 : 0.f;
            const float rz = __frcp_rn(z);
            const float gc0 = gu * rz, gc1 = gv * rz;
            float gc2 = -(gu * u + gv * v) * rz;
            if (p.g_wdepth && c[2] >= 1e-3f) gc2 += p.g_wdepth[o];
            float *gp = p.g_points + (long long)b * 4 * HW + j;
#pragma unroll
            for (int k = 0; k < 4; k++) gp[k * HW] = P[k] * gc0 + P[4 + k] * gc1 + P[8 + k] * gc2;
            gP[0] += gc0 * X0; gP[1] += gc0 * X1; gP[2] += gc0 * X2; gP[3] += gc0 * X3;
            gP[4] += gc1 * X0; gP[5] += gc1 * X1; gP[6] += gc1 * X2; gP[7] += gc1 * X3;
            gP[8] += gc2 * X0; gP[9] += gc2 * X1; gP[10] += gc2 * X2; gP[11] += gc2 * X3;
        }
    }
    if (BWD && p.gP_partial) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int e = 0; e < 12; e++) {
            const float v = warp_sum(gP[e]);
            if (lane == 0) red[wid * 12 + e] = v;
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float t = 0.f;
            for (int w = 0; w < G_NT / 32; w++) t += red[w * 12 + threadIdx.x];
            p.gP_partial[((long long)b * gridDim.x + blockIdx.x) * 12 + threadIdx.x] = t;
        }
    }
}

__global__ void __launch_bounds__(256) reduce_gP2_kernel(const float *partial, int ctas_per_b, float *gP)
{
    __shared__ double sh[8];
    const int b = blockIdx.x / 12, e = blockIdx.x % 12;
    double acc = 0.0;
    for (int i = threadIdx.x; i < ctas_per_b; i += blockDim.x) acc += (double)partial[((long long)b * ctas_per_b + i) * 12 + e];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += sh[i];
        gP[b * 12 + e] = (float)t;
    }
}

// ---- grid_sample ---------------------------------------------------------------------------------
struct GSParams {
    ImgView in;
    const float *grid, *g_out;
    int B, C, H, W, Ho, Wo, border, align;
    float *out, *g_grid;
    ImgViewW g_in;
};

struct GSamp {
    float nw, ne, sw, se, wx, wy, mx, my;
    int x0, y0;
    bool in00, in01, in10, in11;
};

__device__ __forceinline__ float gs_unnormalize(float g, int size, int align, float &mult)
{
    if (align) {            // ((g + 1) / 2) * (size - 1), evaluated as (g + 1) * ((size-1)/2)
        mult = (float)(size - 1) / 2;
        return xmul(xadd(g, 1.0f), mult);
    }
    mult = (float)size / 2;  // ((g + 1) * size - 1) / 2, evaluated as fma(g + 1, size/2, -0.5)
    return xfma(xadd(g, 1.0f), mult, -0.5f);
}

__device__ __forceinline__ void gs_setup(const GSParams &p, float gx, float gy, GSamp &s)
{
    float ix = gs_unnormalize(gx, p.W, p.align, s.mx), iy = gs_unnormalize(gy, p.H, p.align, s.my);
    const float wm1 = (float)(p.W - 1), hm1 = (float)(p.H - 1);
    if (p.border) {
        if (!(ix > 0.0f && ix < wm1)) s.mx = 0.f;
        if (!(iy > 0.0f && iy < hm1)) s.my = 0.f;
        ix = fminf(wm1, fmaxf(0.0f, ix));
        iy = fminf(hm1, fmaxf(0.0f, iy));
    }
    const float xw = floorf(ix), yn = floorf(iy);
    s.wx = xsub(ix, xw);
    s.wy = xsub(iy, yn);
    const float e = xsub(1.0f, s.wx), so = xsub(1.0f, s.wy);
    s.nw = xmul(so, e); s.ne = xmul(so, s.wx); s.sw = xmul(s.wy, e); s.se = xmul(s.wy, s.wx);
    const bool inx0 = (xw >= 0.0f) && (xw <= wm1), inx1 = (xw >= -1.0f) && (xw <= wm1 - 1.0f);
    const bool iny0 = (yn >= 0.0f) && (yn <= hm1), iny1 = (yn >= -1.0f) && (yn <= hm1 - 1.0f);
    s.in00 = iny0 && inx0; s.in01 = iny0 && inx1; s.in10 = iny1 && inx0; s.in11 = iny1 && inx1;
    s.x0 = (inx0 || inx1) ? (int)xw : 0;
    s.y0 = (iny0 || iny1) ? (int)yn : 0;
}

template <bool BWD>
__global__ void __launch_bounds__(G_NT) grid_sample_kernel(const GSParams p)
{
    const long long HWo = (long long)p.Ho * p.Wo, n = (long long)p.B * HWo;
    for (long long i = (long long)blockIdx.x * G_NT + threadIdx.x; i < n; i += (long long)gridDim.x * G_NT) {
        const int b = (int)(i / HWo);
        const long long j = i - (long long)b * HWo;
        GSamp s;
        gs_setup(p, p.grid[i * 2], p.grid[i * 2 + 1], s);
        const long long base = (long long)b * p.in.sb + (long long)s.y0 * p.in.sh + (long long)s.x0 * p.in.sw;
        float gix = 0.f, giy = 0.f;
        for (int c = 0; c < p.C; c++) {
            const float *ip = p.in.p + base + (long long)c * p.in.sc;
            const float v00 = s.in00 ? ip[0] : 0.f, v01 = s.in01 ? ip[p.in.sw] : 0.f;
            const float v10 = s.in10 ? ip[p.in.sh] : 0.f, v11 = s.in11 ? ip[p.in.sh + p.in.sw] : 0.f;
            const long long oo = ((long long)b * p.C + c) * HWo + j;
            if (!BWD) {
                p.out[oo] = xfma(v11, s.se, xfma(v10, s.sw, xfma(v01, s.ne, xmul(v00, s.nw))));
            } else {
                const float g = p.g_out[oo];
                gix += g * ((v01 - v00) * (1.0f - s.wy) + (v11 - v10) * s.wy);
                giy += g * ((v10 - v00) * (1.0f - s.wx) + (v11 - v01) * s.wx);
                if (p.g_in.p) {
                    float *gp = p.g_in.p + (long long)b * p.g_in.sb + (long long)c * p.g_in.sc + (long long)s.y0 * p.g_in.sh + (long long)s.x0 * p.g_in.sw;
                    if (s.in00) atomicAdd(gp, g * s.nw);
                    if (s.in01) atomicAdd(gp + p.g_in.sw, g * s.ne);
                    if (s.in10) atomicAdd(gp + p.g_in.sh, g * s.sw);
                    if (s.in11) atomicAdd(gp + p.g_in.sh + p.g_in.sw, g * s.se);
                }
            }
        }
        if (BWD && p.g_grid) {
            p.g_grid[i * 2] = gix * s.mx;
            p.g_grid[i * 2 + 1] = giy * s.my;
        }
    }
}

}  // namespace e2e

using namespace e2e;

extern "C" {

int e2e_backproject_fwd(const float *depth, const float *inv_K, int B, int H, int W, float *cam_points, void *stream)
{
    E2E_REQUIRE(depth && inv_K && cam_points && B > 0 && H > 0 && W > 0, "backproject: bad arguments");
    backproject_fwd_kernel<<<ew_grid((long long)B * H * W), G_NT, 0, (cudaStream_t)stream>>>(depth, inv_K, B, H, W, cam_points);
    count_launch();
    return finish_launch("backproject_fwd");
}

int e2e_backproject_bwd(const float *grad_cam, const float *inv_K, int B, int H, int W, float *grad_depth, void *stream)
{
    E2E_REQUIRE(grad_cam && inv_K && grad_depth && B > 0 && H > 0 && W > 0, "backproject_bwd: bad arguments");
    backproject_bwd_kernel<<<ew_grid((long long)B * H * W), G_NT, 0, (cudaStream_t)stream>>>(grad_cam, inv_K, B, H, W, grad_depth);
    count_launch();
    return finish_launch("backproject_bwd");
}

static int p3_fill(P3Params &p, const float *points, const float *K, const float *T, int B, int H, int W, float eps, cudaStream_t st)
{
    E2E_REQUIRE(points && K && T && B > 0 && B <= 65535 && H >= 2 && W >= 2, "project3d: bad arguments");
    p.points = points; p.K = K; p.T = T; p.B = B; p.H = H; p.W = W; p.eps = eps;
    p.wm1 = (float)(W - 1); p.hm1 = (float)(H - 1);
    const DivC dW = host_divc(p.wm1, st), dH = host_divc(p.hm1, st);
    p.rcpW = dW.rcp; p.rcpH = dH.rcp; p.exact = dW.exact && dH.exact;
    return 0;
}

static inline int p3_blocks(int HW)
{
    int b = (HW + G_NT - 1) / G_NT;
    return b > kNumSMs * 2 ? kNumSMs * 2 : b;
}

int e2e_project3d_fwd(const float *points, const float *K, const float *T, int B, int H, int W, float eps,
                      float *pix, float *valid, float *warped_depth, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    P3Params p = {};
    if (int rc = p3_fill(p, points, K, T, B, H, W, eps, st)) return rc;
    E2E_REQUIRE(pix && valid, "project3d: null output");
    p.pix = pix; p.valid = valid; p.wdepth = warped_depth;
    project3d_kernel<false><<<dim3(p3_blocks(H * W), B), G_NT, 0, st>>>(p);
    count_launch();
    return finish_launch("project3d_fwd");
}

int e2e_project3d_bwd(const float *points, const float *K, const float *T, int B, int H, int W, float eps,
                      const float *grad_pix, const float *grad_warped_depth,
                      float *grad_points, float *grad_P, void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    P3Params p = {};
    if (int rc = p3_fill(p, points, K, T, B, H, W, eps, st)) return rc;
    E2E_REQUIRE(grad_points && (grad_pix || grad_warped_depth), "project3d_bwd: null gradient");
    p.g_pix = grad_pix; p.g_wdepth = grad_warped_depth; p.g_points = grad_points;
    const int blocks = p3_blocks(H * W);
    if (grad_P) {
        E2E_REQUIRE(workspace && workspace_bytes >= sizeof(float) * 12 * (size_t)blocks * B, "project3d_bwd: workspace too small");
        p.gP_partial = (float *)workspace;
    }
    project3d_kernel<true><<<dim3(blocks, B), G_NT, 0, st>>>(p);
    count_launch();
    if (int rc = finish_launch("project3d_bwd")) return rc;
    if (grad_P) {
        reduce_gP2_kernel<<<B * 12, 256, 0, st>>>(p.gP_partial, blocks, grad_P);
        count_launch();
        return finish_launch("reduce_gP2");
    }
    return 0;
}

static int gs_fill(GSParams &p, const float *input, const int64_t in_strides[4], const float *grid,
                   int B, int C, int H, int W, int Ho, int Wo, int padding_mode, int align_corners)
{
    E2E_REQUIRE(input && grid && B > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "grid_sample: bad arguments");
    E2E_REQUIRE(padding_mode == 0 || padding_mode == 1, "grid_sample: padding_mode must be 0 (zeros) or 1 (border)");
    p.in = make_view(input, in_strides); p.grid = grid;
    p.B = B; p.C = C; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.border = padding_mode; p.align = align_corners ? 1 : 0;
    return 0;
}

int e2e_grid_sample_fwd(const float *input, const int64_t in_strides[4], const float *grid,
                        int B, int C, int H, int W, int Ho, int Wo, int padding_mode, int align_corners,
                        float *output, void *stream)
{
    GSParams p = {};
    if (int rc = gs_fill(p, input, in_strides, grid, B, C, H, W, Ho, Wo, padding_mode, align_corners)) return rc;
    E2E_REQUIRE(output, "grid_sample: null output");
    p.out = output;
    grid_sample_kernel<false><<<ew_grid((long long)B * Ho * Wo), G_NT, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return finish_launch("grid_sample_fwd");
}

int e2e_grid_sample_bwd(const float *grad_output, const float *input, const int64_t in_strides[4],
                        const float *grid, int B, int C, int H, int W, int Ho, int Wo,
                        int padding_mode, int align_corners,
                        float *grad_input, const int64_t grad_in_strides[4], float *grad_grid, void *stream)
{
    GSParams p = {};
    if (int rc = gs_fill(p, input, in_strides, grid, B, C, H, W, Ho, Wo, padding_mode, align_corners)) return rc;
    E2E_REQUIRE(grad_output && (grad_input || grad_grid), "grid_sample_bwd: nothing to compute");
    p.g_out = grad_output; p.g_grid = grad_grid;
    if (grad_input) {
        E2E_REQUIRE(grad_in_strides, "grid_sample_bwd: grad_input needs strides");
        p.g_in = make_view_w(grad_input, grad_in_strides);
    }
    grid_sample_kernel<true><<<ew_grid((long long)B * Ho * Wo), G_NT, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return finish_launch("grid_sample_bwd");
}

}  // extern "C"
