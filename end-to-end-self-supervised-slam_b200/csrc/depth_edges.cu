// The elementwise passes on either side of the depth network, each fused into one kernel forward
// and one backward:
//   disp -> depth:   depth = (1 / disp) * ratio          online_adaption.py:282, 295-298; train_depth.py:323-340
//                    (reciprocal and scaling are two roundings, as in the reference; ratio = median(gt) / median(depth) is a
//                    device scalar, NULL = no scaling)
//   dual disparity:  process_disparity (train_depth.py:224-237): blend of the disparity of the frame and of its mirror image,
//                    out = m * left + m * right' + (1 - m - m) * 0.5 (left + right'),  right' = right flipped along W,
//                    m = the row mask 1 - clip(20 (linspace(0,1,H) - 0.05), 0, 1) (passed in, built by torch: H values)
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

}  // extern "C"
