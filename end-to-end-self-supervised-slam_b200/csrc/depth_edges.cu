// The elementwise passes on either side of the depth network, each fused into one kernel forward
// and one backward:
//   disp -> depth:   depth = (1 / disp) * ratio          online_adaption.py:282, 295-298; train_depth.py:323-340
//                    (reciprocal and scaling are two roundings, as in the reference; ratio = median(gt) / median(depth) is a
//                    device scalar, NULL = no scaling)
//   dual disparity:  process_disparity (train_depth.py:224-237): blend of the disparity of the frame and of its mirror image,
//                    out = m * left + m * right' + (1 - m - m) * 0.5 (left + right'),  right' = right flipped along W,
//                    m = the row mask 1 - clip(20 (linspace(0,1,H) - 0.05), 0, 1) (passed in, built by torch: H values)
//   median:          torch.median (online_adaption.py:295: ratio = median(gt) / median(depth)) as a k-th order statistic by four rounds of
//                    8-bit radix select over an order-preserving integer image of the floats -- no sort, no host synchronisation
#include "common.cuh"

namespace e2e {

__global__ void __launch_bounds__(256) disp_to_depth_fwd_kernel(const float *disp, const float *ratio, long long n, float *depth)
{
    const bool scaled = ratio != nullptr;
    const float r = scaled ? __ldg(ratio) : 1.0f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float d = xdiv(1.0f, disp[i]);
        depth[i] = scaled ? xmul(d, r) : d;
    }
}

// d depth / d disp = -ratio / disp^2
__global__ void __launch_bounds__(256) disp_to_depth_bwd_kernel(const float *disp, const float *ratio, const float *g_depth, long long n,
                                                                float *g_disp)
{
    const float r = ratio ? __ldg(ratio) : 1.0f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float d = disp[i];
        g_disp[i] = -g_depth[i] * r / (d * d);
    }
}

__global__ void __launch_bounds__(256) dual_disp_fwd_kernel(const float *left, const float *right, const float *row_mask, int H, int W,
                                                            float *out)
{
    for (int i = blockIdx.x * 256 + threadIdx.x; i < H * W; i += gridDim.x * 256) {
        const int y = i / W, x = i - y * W;
        const float m = row_mask[y], l = left[i], r = right[y * W + (W - 1 - x)];
        const float mid = xmul(0.5f, xadd(l, r));
        // r_mask * left + l_mask * right + (1.0 - l_mask - r_mask) * middle, left to right (r_mask == l_mask: the mask does not vary along W)
        out[i] = xadd(xadd(xmul(m, l), xmul(m, r)), xmul(xsub(xsub(1.0f, m), m), mid));
    }
}

__global__ void __launch_bounds__(256) dual_disp_bwd_kernel(const float *g_out, const float *row_mask, int H, int W, float *g_left, float *g_right)
{
    for (int i = blockIdx.x * 256 + threadIdx.x; i < H * W; i += gridDim.x * 256) {
        const int y = i / W, x = i - y * W;
        const float m = row_mask[y];
        const float w = m + 0.5f * (1.0f - m - m);        // weight of either input at a pixel
        g_left[i] = g_out[i] * w;
        g_right[i] = g_out[y * W + (W - 1 - x)] * w;     // right[y, x] feeds out[y, W-1-x]
    }
}

// ---- k-th smallest element (0-based) of an fp32 array: radix select, 4 rounds x (histogram of the next 8 key bits among the
// elements that match the prefix found so far | pick the bin that holds rank k) ------------------------------------------------------
struct SelectState {
    unsigned prefix;               // key bits decided so far (high bits)
    unsigned nan;                  // a NaN was seen: torch.median then returns NaN
    unsigned long long k;          // rank still to be found inside the current prefix
    unsigned hist[256];
};

__device__ __forceinline__ unsigned float_key(float f)          // unsigned order == float order (-inf ... -0 +0 ... +inf)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void select_init_kernel(SelectState *st, unsigned long long k)
{
    if (threadIdx.x == 0) { st->prefix = 0u; st->nan = 0u; st->k = k; }
    st->hist[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) select_hist_kernel(const float *x, long long n, int round, SelectState *st)
{
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const int shift = 24 - 8 * round;
    const unsigned himask = round ? (0xffffffffu << (shift + 8)) : 0u, prefix = st->prefix;
    bool nan = false;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float v = x[i];
        nan |= (v != v);
        const unsigned key = float_key(v);
        if ((key & himask) == prefix) atomicAdd(&sh[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
    if (round == 0 && nan) st->nan = 1u;
}

__global__ void select_pick_kernel(SelectState *st, int round, float *out)
{
    __shared__ unsigned h[256];
    h[threadIdx.x] = st->hist[threadIdx.x];
    st->hist[threadIdx.x] = 0u;                                 // ready for the next round
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long k = st->k, cum = 0;
        int b = 0;
        for (; b < 255; b++) {
            if (cum + h[b] > k) break;
            cum += h[b];
        }
        const int shift = 24 - 8 * round;
        const unsigned prefix = st->prefix | ((unsigned)b << shift);
        st->prefix = prefix;
        st->k = k - cum;
        if (round == 3) {
            const unsigned u = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix;
            out[0] = st->nan ? __int_as_float(0x7fc00000) : __uint_as_float(u);
        }
    }
}

static int de_blocks(long long n)
{
    long long b = (n + 255) / 256;
    if (b > kNumSMs * 16) b = kNumSMs * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace e2e

using namespace e2e;

extern "C" {

int e2e_disp_to_depth_fwd(const float *disp, const float *ratio, long long n, float *depth, void *stream)
{
    E2E_REQUIRE(disp && depth && n > 0, "disp_to_depth: bad arguments");
    disp_to_depth_fwd_kernel<<<de_blocks(n), 256, 0, (cudaStream_t)stream>>>(disp, ratio, n, depth);
    count_launch();
    return finish_launch("disp_to_depth_fwd");
}

int e2e_disp_to_depth_bwd(const float *disp, const float *ratio, const float *grad_depth, long long n, float *grad_disp, void *stream)
{
    E2E_REQUIRE(disp && grad_depth && grad_disp && n > 0, "disp_to_depth_bwd: bad arguments");
    disp_to_depth_bwd_kernel<<<de_blocks(n), 256, 0, (cudaStream_t)stream>>>(disp, ratio, grad_depth, n, grad_disp);
    count_launch();
    return finish_launch("disp_to_depth_bwd");
}

int e2e_dual_disparity_fwd(const float *left, const float *right, const float *row_mask, int H, int W, float *out, void *stream)
{
    E2E_REQUIRE(left && right && row_mask && out && H > 0 && W > 0, "dual_disparity: bad arguments");
    dual_disp_fwd_kernel<<<de_blocks((long long)H * W), 256, 0, (cudaStream_t)stream>>>(left, right, row_mask, H, W, out);
    count_launch();
    return finish_launch("dual_disparity_fwd");
}

int e2e_dual_disparity_bwd(const float *grad_out, const float *row_mask, int H, int W, float *grad_left, float *grad_right, void *stream)
{
    E2E_REQUIRE(grad_out && row_mask && grad_left && grad_right && H > 0 && W > 0, "dual_disparity_bwd: bad arguments");
    dual_disp_bwd_kernel<<<de_blocks((long long)H * W), 256, 0, (cudaStream_t)stream>>>(grad_out, row_mask, H, W, grad_left, grad_right);
    count_launch();
    return finish_launch("dual_disparity_bwd");
}

size_t e2e_select_workspace_bytes(void) { return sizeof(SelectState) + 256; }

int e2e_select_kth(const float *x, long long n, long long k, float *out, void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(x && out && n > 0 && k >= 0 && k < n, "select_kth: bad arguments (n=%lld, k=%lld)", n, k);
    E2E_REQUIRE(workspace && workspace_bytes >= sizeof(SelectState), "select_kth: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    SelectState *s = (SelectState *)workspace;
    select_init_kernel<<<1, 256, 0, st>>>(s, (unsigned long long)k);
    for (int round = 0; round < 4; round++) {
        select_hist_kernel<<<de_blocks(n), 256, 0, st>>>(x, n, round, s);
        select_pick_kernel<<<1, 256, 0, st>>>(s, round, out);
    }
    count_launch(9);
    return finish_launch("select_kth");
}

}  // extern "C"
