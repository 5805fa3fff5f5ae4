// Edge-aware disparity smoothness (train_depth.py:763-773 + loss/losses.py:119-132), value AND gradient in one sweep.
//
//   loss = mean_x |n_i - n_right| exp(-mean_c |I_i - I_right|) + mean_y |n_i - n_down| exp(-mean_c |I_i - I_down|),
//   n = disp / (mean_hw(disp) + 1e-7)
//
// The separate forward / backward entry points of small_losses.cu evaluate every pair's edge weight from both of its
// pixels, in two backward passes, with per-CTA re-reduction of the mean: 4.7 ms per 256 x 480 x 640 (10 % of HBM peak).
// Here every thread walks down one image column: the pair with the row above is formed from values carried in
// registers, the pair with the right neighbour from the neighbouring lane's registers (shuffles), so each pixel is
// loaded once, each edge weight is computed once, and d loss / d n is written in the same pass:
//   smooth_mean (per-image sum)  ->  smooth_me (mean + 1e-7)  ->  smooth_vg (loss partials, gn = dL/dn, dot = sum gn * disp)
//   ->  smooth_stats (per image)  ->  smooth_loss;   backward = one elementwise pass  grad = up * (gn / me - dot / (me^2 HW)).
// Compulsory traffic 4 (mean) + 16 (vg read) + 4 (gn write) + 8 (backward) = 32 B/px.
#include <cstdint>

#include "common.cuh"

namespace e2e {

constexpr int SV_NT = 128;          // threads per CTA
constexpr int SV_OWN = 30;          // owner columns per warp: lanes 1..30; lanes 0 and 31 carry the left / right neighbour column
constexpr int SV_COLS = (SV_NT / 32) * SV_OWN;      // owner columns per CTA
constexpr int SV_ROWS = 30;         // rows per CTA (480 = 16 x 30)

__global__ void __launch_bounds__(256) smooth_sum_kernel(const float *disp, int HW, double *partial)
{
    __shared__ double sh[8];
    const float *d = disp + (long long)blockIdx.y * HW;
    double v = 0.0;
    if ((HW & 3) == 0 && (((uintptr_t)d) & 15u) == 0) {      // 16-byte loads, four per thread in flight
        const float4 *d4 = reinterpret_cast<const float4 *>(d);
        for (int i = blockIdx.x * 256 + threadIdx.x; i < HW / 4; i += gridDim.x * 256) {
            const float4 q = d4[i];
            v += ((double)q.x + (double)q.y) + ((double)q.z + (double)q.w);
        }
    } else {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) v += (double)d[i];
    }
    v = warp_sum_d(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// me[b] = (float)(sum / HW) + 1e-7 (train_depth.py:769), one warp per image, fixed order
__global__ void __launch_bounds__(32) smooth_me_kernel(const double *partial, int nblk, int HW, float *me)
{
    double t = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 32) t += partial[(long long)blockIdx.x * nblk + i];
    t = warp_sum_d(t);
    if (threadIdx.x == 0) me[blockIdx.x] = xadd((float)(t / (double)HW), 1e-7f);
}

struct SmoothVG {
    const float *disp;
    ImgView img;
    int B, H, W;
    const float *me;          // [B]
    float cx, cy;             // 1 / (B H (W-1)),  1 / (B (H-1) W)
    float *gn;                // [B,H,W]  d loss / d n for an upstream gradient of 1
    double *partial;          // [B][ctas per image][3] = {sum of x terms, sum of y terms, sum of gn * disp}
};

struct SPix {
    float n, c0, c1, c2;
};

// exp(-mean_c |a - b|): channels summed in order (losses.py:125-126).  The weight is a smooth factor of the loss (contract: 1e-5),
// so the mean is a multiplication by 1/3 and the exponential the hardware ex2 (2 ulp): an IEEE division and a libm expf per pair were
// a fifth of this kernel's instructions.  What has to agree with the reference bit for bit is the SIGN of the disparity difference
// (the gradient flips with it), and that comes from the exactly divided n.
__device__ __forceinline__ float edge_weight(const SPix &a, const SPix &b)
{
    return __expf(-(((fabsf(a.c0 - b.c0) + fabsf(a.c1 - b.c1)) + fabsf(a.c2 - b.c2)) * (1.0f / 3.0f)));
}

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(SV_NT) smooth_vg_kernel(const SmoothVG p)
{
    __shared__ double sh[SV_NT / 32][3];
    const int b = blockIdx.z, H = p.H, W = p.W;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // a warp owns 30 columns and also loads the column on either side of them (6.7 % more loads): every horizontal pair is then formed
    // from shuffles alone -- the version in which lanes 0 / 31 fetched and normalised their outside neighbour themselves spent a
    // fifth of the warp's instructions on those two lanes
    const int x = blockIdx.x * SV_COLS + wid * SV_OWN + lane - 1;
    const int y0 = blockIdx.y * SV_ROWS, y1 = min(y0 + SV_ROWS, H);
    const bool col = x >= 0 && x < W;
    const bool own = col && lane >= 1 && lane <= SV_OWN;
    const float *d = p.disp + (long long)b * H * W;
    const long long ib = (long long)b * p.img.sb;
    const float me = p.me[b];
    float sx = 0.f, sy = 0.f, dot = 0.f;
    SPix up = {0.f, 0.f, 0.f, 0.f};
    float g_row = 0.f;        // x-pair part of gn of the previous row's pixel, minus the y pair above it
    float d_prev = 0.f;
    // rows y0-1 (only to seed the pair with the first owned row) .. y1 (only to close the pair below the last owned row);
    // the next row's four values are requested before the current row is worked on
    const int y_first = max(y0 - 1, 0), y_last = min(y1, H - 1);
    float nd = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
    if (col) {
        const long long o = ib + (long long)y_first * p.img.sh + (long long)x * p.img.sw;
        nd = d[y_first * W + x]; n0 = p.img.p[o]; n1 = p.img.p[o + p.img.sc]; n2 = p.img.p[o + 2 * p.img.sc];
    }
    for (int y = y_first; y <= y_last; y++) {
        const float dv = nd;
        SPix c = {0.f, n0, n1, n2};
        if (col && y < y_last) {
            const long long o = ib + (long long)(y + 1) * p.img.sh + (long long)x * p.img.sw;
            nd = d[(y + 1) * W + x]; n0 = p.img.p[o]; n1 = p.img.p[o + p.img.sc]; n2 = p.img.p[o + 2 * p.img.sc];
        }
        if (col) c.n = xdiv(dv, me);
        const bool owned = y >= y0 && y < y1;
        float gx = 0.f, gxl = 0.f;
        if (owned) {      // uniform per CTA
            SPix r;       // right neighbour: the next lane's pixel
            r.n = __shfl_down_sync(0xffffffffu, c.n, 1);
            r.c0 = __shfl_down_sync(0xffffffffu, c.c0, 1);
            r.c1 = __shfl_down_sync(0xffffffffu, c.c1, 1);
            r.c2 = __shfl_down_sync(0xffffffffu, c.c2, 1);
            if (col && lane < 31 && x + 1 < W) {
                const float e = edge_weight(c, r), df = xsub(c.n, r.n);      // explicitly rounded: both pixels of a pair agree on the sign
                if (own) sx += fabsf(df) * e;                                   // a pair is counted by its left pixel
                gx = p.cx * sgnf(df) * e;
            }
            gxl = __shfl_up_sync(0xffffffffu, gx, 1);      // the pair with the left neighbour (lane 0's own left pair is nobody's business here)
        }
        // the pair (y-1, y): closes gn of row y-1
        float gy = 0.f;
        if (own && y > max(y0 - 1, 0)) {
            const float e = edge_weight(up, c), df = xsub(up.n, c.n);
            gy = p.cy * sgnf(df) * e;
            if (y - 1 >= y0) {                       // the pair belongs to the segment that owns its upper pixel
                sy += fabsf(df) * e;
                const float g = g_row + gy;
                p.gn[((long long)b * H + (y - 1)) * W + x] = g;
                dot += g * d_prev;
            }
        }
        g_row = gx - gxl - gy;
        up = c;
        d_prev = dv;
    }
    if (own && y1 == H) {                            // the last image row has no pair below it
        p.gn[((long long)b * H + (H - 1)) * W + x] = g_row;
        dot += g_row * d_prev;
    }
    double v[3] = {(double)sx, (double)sy, (double)dot};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        v[k] = warp_sum_d(v[k]);
        if (lane == 0) sh[wid][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < SV_NT / 32; w++) t += sh[w][threadIdx.x];
        const long long cta = ((long long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.partial[cta * 3 + threadIdx.x] = t;
    }
}

// per image: {x sum, y sum} and the backward scalars inv = 1 / me, corr = dot / (me^2 HW); one warp per image, fixed order
// (raw = the disparity was not normalised here: inv = 1, no correction term)
__global__ void __launch_bounds__(32) smooth_stats_kernel(const double *partial, int ctas, const float *me, int HW, double *sums, float *stats, int raw)
{
    double t[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < ctas; i += 32)
#pragma unroll
        for (int k = 0; k < 3; k++) t[k] += partial[((long long)blockIdx.x * ctas + i) * 3 + k];
#pragma unroll
    for (int k = 0; k < 3; k++) t[k] = warp_sum_d(t[k]);
    if (threadIdx.x == 0) {
        const float m = me[blockIdx.x];
        sums[blockIdx.x * 2] = t[0];
        sums[blockIdx.x * 2 + 1] = t[1];
        stats[blockIdx.x * 2] = raw ? 1.0f : 1.0f / m;
        stats[blockIdx.x * 2 + 1] = raw ? 0.0f : (float)(t[2] / ((double)m * (double)m * (double)HW));
    }
}

__global__ void __launch_bounds__(32) smooth_loss_kernel(const double *sums, int B, double inv_nx, double inv_ny, float *loss)
{
    double tx = 0.0, ty = 0.0;
    for (int i = threadIdx.x; i < B; i += 32) { tx += sums[i * 2]; ty += sums[i * 2 + 1]; }
    tx = warp_sum_d(tx); ty = warp_sum_d(ty);
    if (threadIdx.x == 0) loss[0] = (float)(tx * inv_nx + ty * inv_ny);
}

__global__ void smooth_ones_kernel(float *me, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) me[i] = 1.0f;          // x / 1 is x: the sweep then works on the disparity as given
}

__global__ void __launch_bounds__(256) smooth_apply_kernel(const float *gn, const float *stats, const float *grad_loss, int HW, float *grad_disp)
{
    const int b = blockIdx.y;
    const float up = grad_loss ? __ldg(grad_loss) : 1.0f, inv = stats[b * 2], corr = stats[b * 2 + 1];
    const float *g = gn + (long long)b * HW;
    float *o = grad_disp + (long long)b * HW;
    if ((HW & 3) == 0 && ((((uintptr_t)g) | ((uintptr_t)o)) & 15u) == 0) {
        const float4 *g4 = reinterpret_cast<const float4 *>(g);
        float4 *o4 = reinterpret_cast<float4 *>(o);
        for (int i = blockIdx.x * 256 + threadIdx.x; i < HW / 4; i += gridDim.x * 256) {
            const float4 q = g4[i];
            o4[i] = make_float4(up * (q.x * inv - corr), up * (q.y * inv - corr), up * (q.z * inv - corr), up * (q.w * inv - corr));
        }
        return;
    }
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) o[i] = up * (g[i] * inv - corr);
}

static size_t sv_a256(size_t n) { return (n + 255) / 256 * 256; }
static int sv_sum_blocks(int HW)
{
    int b = (HW + 256 * 8 - 1) / (256 * 8);
    return b < 1 ? 1 : (b > 256 ? 256 : b);
}

}  // namespace e2e

using namespace e2e;

extern "C" {

size_t e2e_smooth_vg_workspace_bytes(int B, int H, int W)
{
    if (B < 1 || H < 1 || W < 1) return 256;
    const size_t ctas = (size_t)((W + SV_COLS - 1) / SV_COLS) * ((H + SV_ROWS - 1) / SV_ROWS);
    return sv_a256((size_t)B * sv_sum_blocks(H * W) * 8) + sv_a256((size_t)B * 4) + sv_a256((size_t)B * ctas * 3 * 8) + sv_a256((size_t)B * 2 * 8) + 256;
}

static int smooth_vg_run(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                         float *loss, float *gn, float *stats, void *workspace, size_t workspace_bytes, void *stream, int raw)
{
    cudaStream_t st = (cudaStream_t)stream;
    E2E_REQUIRE(disp && img && img_strides && loss && gn && stats && workspace, "smooth_vg: null argument");
    E2E_REQUIRE(B >= 1 && H >= 2 && W >= 2 && B <= 65535, "smooth_vg: needs H, W >= 2 (the means over H*(W-1) and (H-1)*W pairs)");
    E2E_REQUIRE(workspace_bytes >= e2e_smooth_vg_workspace_bytes(B, H, W), "smooth_vg: workspace too small");
    const int HW = H * W, nsum = sv_sum_blocks(HW);
    const dim3 grid((W + SV_COLS - 1) / SV_COLS, (H + SV_ROWS - 1) / SV_ROWS, B);
    E2E_REQUIRE(grid.y <= 65535, "smooth_vg: image too tall");
    const int ctas = (int)(grid.x * grid.y);
    unsigned char *w = (unsigned char *)workspace;
    double *sum_partial = (double *)w;      w += sv_a256((size_t)B * nsum * 8);
    float *me = (float *)w;                 w += sv_a256((size_t)B * 4);
    double *partial = (double *)w;          w += sv_a256((size_t)B * ctas * 3 * 8);
    double *sums = (double *)w;
    SmoothVG p;
    p.disp = disp; p.img = make_view(img, img_strides); p.B = B; p.H = H; p.W = W; p.me = me;
    p.cx = (float)(1.0 / ((double)B * H * (W - 1))); p.cy = (float)(1.0 / ((double)B * (H - 1) * W));
    p.gn = gn; p.partial = partial;
    if (raw) {
        smooth_ones_kernel<<<(B + 255) / 256, 256, 0, st>>>(me, B);
    } else {
        smooth_sum_kernel<<<dim3(nsum, B), 256, 0, st>>>(disp, HW, sum_partial);
        smooth_me_kernel<<<B, 32, 0, st>>>(sum_partial, nsum, HW, me);
    }
    smooth_vg_kernel<<<grid, SV_NT, 0, st>>>(p);
    smooth_stats_kernel<<<B, 32, 0, st>>>(partial, ctas, me, HW, sums, stats, raw);
    smooth_loss_kernel<<<1, 32, 0, st>>>(sums, B, 1.0 / ((double)B * H * (W - 1)), 1.0 / ((double)B * (H - 1) * W), loss);
    count_launch(raw ? 4 : 5);
    return finish_launch("smooth_vg");
}

int e2e_smooth_vg(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                  float *loss, float *gn, float *stats, void *workspace, size_t workspace_bytes, void *stream)
{
    return smooth_vg_run(disp, img, img_strides, B, H, W, loss, gn, stats, workspace, workspace_bytes, stream, 0);
}

int e2e_smooth_vg_raw(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                      float *loss, float *gn, float *stats, void *workspace, size_t workspace_bytes, void *stream)
{
    return smooth_vg_run(disp, img, img_strides, B, H, W, loss, gn, stats, workspace, workspace_bytes, stream, 1);
}

int e2e_smooth_apply(const float *gn, const float *stats, const float *grad_loss, int B, int H, int W, float *grad_disp, void *stream)
{
    E2E_REQUIRE(gn && stats && grad_disp && B >= 1 && B <= 65535 && H >= 1 && W >= 1, "smooth_apply: bad arguments");
    const int HW = H * W;
    int blocks = (HW + 256 * 4 - 1) / (256 * 4);
    if (blocks > 1024) blocks = 1024;
    smooth_apply_kernel<<<dim3(blocks, B), 256, 0, (cudaStream_t)stream>>>(gn, stats, grad_loss, HW, grad_disp);
    count_launch();
    return finish_launch("smooth_apply");
}

}  // extern "C"
