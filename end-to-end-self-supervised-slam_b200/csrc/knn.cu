// Exact K = 1 nearest neighbour (chamferdist.chamfer.knn_points as used by loss/losses.py:39-63 and
// online_adaption.py:638-645), with gradslam's transform_pointcloud optionally fused into the query load.
//
// Brute force, tiled through shared memory: every thread owns QPT query points in registers and sweeps
// the reference cloud in tiles that the CTA stages cooperatively; all lanes read the same reference point
// (shared-memory broadcast).  Squared distances are evaluated in the oracle's operation order
// ((dx*dx + dy*dy) + dz*dz, one rounding each) and ties keep the lowest index, so `idx` is bit-exact.
// This kernel is FP32-issue bound (about 11 instructions per point pair), not HBM bound.
#include "common.cuh"

namespace e2e {

constexpr int KNN_NT = 256;
constexpr int KNN_QPT = 4;
constexpr int KNN_TILE = 2048;

__device__ __forceinline__ void load_query(const float *query, const float *T, long long i, float q[3])
{
    const float x = query[i * 3], y = query[i * 3 + 1], z = query[i * 3 + 2];
    if (T) {   // R p + t, accumulated left to right (oracle/fusion_oracle.py: transform_pointcloud)
#pragma unroll
        for (int r = 0; r < 3; r++)
            q[r] = xadd(xadd(xadd(xmul(T[r * 4], x), xmul(T[r * 4 + 1], y)), xmul(T[r * 4 + 2], z)), T[r * 4 + 3]);
    } else {
        q[0] = x; q[1] = y; q[2] = z;
    }
}

__global__ void __launch_bounds__(KNN_NT) knn1_fwd_kernel(const float *query, const float *T, const float *ref, long long P1, long long P2,
                                                          float *dist2, long long *idx)
{
    __shared__ float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    float q[KNN_QPT][3], best[KNN_QPT];
    long long bi[KNN_QPT];
    const long long q0 = ((long long)blockIdx.x * KNN_NT + threadIdx.x) * KNN_QPT;
#pragma unroll
    for (int k = 0; k < KNN_QPT; k++) {
        const long long i = q0 + k;
        if (i < P1) load_query(query, T, i, q[k]);
        else q[k][0] = q[k][1] = q[k][2] = 0.f;
        best[k] = INFINITY;
        bi[k] = 0;
    }
    for (long long t0 = 0; t0 < P2; t0 += KNN_TILE) {
        const int tn = (int)((P2 - t0 < KNN_TILE) ? (P2 - t0) : KNN_TILE);
        __syncthreads();
        for (int j = threadIdx.x; j < tn; j += KNN_NT) {
            sx[j] = ref[(t0 + j) * 3];
            sy[j] = ref[(t0 + j) * 3 + 1];
            sz[j] = ref[(t0 + j) * 3 + 2];
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < tn; j++) {
            const float rx = sx[j], ry = sy[j], rz = sz[j];
#pragma unroll
            for (int k = 0; k < KNN_QPT; k++) {
                const float dx = xsub(q[k][0], rx), dy = xsub(q[k][1], ry), dz = xsub(q[k][2], rz);
                const float d2 = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
                if (d2 < best[k]) { best[k] = d2; bi[k] = t0 + j; }     // strict: first minimum wins
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KNN_QPT; k++) {
        const long long i = q0 + k;
        if (i < P1) { dist2[i] = best[k]; idx[i] = bi[k]; }
    }
}

__global__ void __launch_bounds__(KNN_NT) knn1_bwd_kernel(const float *query, const float *T, const float *ref, long long P1,
                                                          const long long *idx, const float *g, float *gq, float *gr)
{
    for (long long i = (long long)blockIdx.x * KNN_NT + threadIdx.x; i < P1; i += (long long)gridDim.x * KNN_NT) {
        float q[3];
        load_query(query, T, i, q);
        const long long j = idx[i];
        const float gi = 2.0f * g[i];
        const float d0 = gi * (q[0] - ref[j * 3]), d1 = gi * (q[1] - ref[j * 3 + 1]), d2 = gi * (q[2] - ref[j * 3 + 2]);
        if (gq) {
            if (T) {   // back through R p + t
                gq[i * 3] = T[0] * d0 + T[4] * d1 + T[8] * d2;
                gq[i * 3 + 1] = T[1] * d0 + T[5] * d1 + T[9] * d2;
                gq[i * 3 + 2] = T[2] * d0 + T[6] * d1 + T[10] * d2;
            } else {
                gq[i * 3] = d0; gq[i * 3 + 1] = d1; gq[i * 3 + 2] = d2;
            }
        }
        if (gr) {
            atomicAdd(gr + j * 3, -d0);
            atomicAdd(gr + j * 3 + 1, -d1);
            atomicAdd(gr + j * 3 + 2, -d2);
        }
    }
}


// gradslam.geometry.geometryutils.transform_pointcloud (online_adaption.py:642): R p + t for (N,3) points, accumulated left to
// right like the oracle (and like load_query above, so that transform-then-kNN equals the fused query load bit for bit).
__global__ void __launch_bounds__(KNN_NT) transform_points_fwd_kernel(const float *pts, const float *T, long long n, float *out)
{
    for (long long i = (long long)blockIdx.x * KNN_NT + threadIdx.x; i < n; i += (long long)gridDim.x * KNN_NT) {
        float q[3];
        load_query(pts, T, i, q);
        out[i * 3] = q[0]; out[i * 3 + 1] = q[1]; out[i * 3 + 2] = q[2];
    }
}

__global__ void __launch_bounds__(KNN_NT) transform_points_bwd_kernel(const float *T, const float *g, long long n, float *gp)
{
    for (long long i = (long long)blockIdx.x * KNN_NT + threadIdx.x; i < n; i += (long long)gridDim.x * KNN_NT) {
        const float d0 = g[i * 3], d1 = g[i * 3 + 1], d2 = g[i * 3 + 2];
        gp[i * 3] = T[0] * d0 + T[4] * d1 + T[8] * d2;              // R^T g
        gp[i * 3 + 1] = T[1] * d0 + T[5] * d1 + T[9] * d2;
        gp[i * 3 + 2] = T[2] * d0 + T[6] * d1 + T[10] * d2;
    }
}

// color_points_loss (loss/losses.py:65-82): mean | noisy_col[i] - gt_col[idx[i]] | over the 3 P1 colour values; per-CTA partial
// sums, reduced in fixed order by the caller's second launch (deterministic).
__global__ void __launch_bounds__(KNN_NT) color_points_fwd_kernel(const float *gt, const float *noisy, const long long *idx, long long P1,
                                                                  float *partial)
{
    __shared__ float scratch[KNN_NT / 32];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * KNN_NT + threadIdx.x; i < P1; i += (long long)gridDim.x * KNN_NT) {
        const long long j = idx[i];
#pragma unroll
        for (int c = 0; c < 3; c++) acc += fabsf(noisy[i * 3 + c] - gt[j * 3 + c]);
    }
    const float t = block_sum(acc, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(KNN_NT) color_points_bwd_kernel(const float *gt, const float *noisy, const long long *idx, long long P1,
                                                                  const float *g, float scale, float *g_noisy, float *g_gt)
{
    const float gs = scale * (g ? __ldg(g) : 1.0f);
    for (long long i = (long long)blockIdx.x * KNN_NT + threadIdx.x; i < P1; i += (long long)gridDim.x * KNN_NT) {
        const long long j = idx[i];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float d = noisy[i * 3 + c] - gt[j * 3 + c];
            const float s = (d > 0.f) ? gs : ((d < 0.f) ? -gs : 0.f);       // torch: sign(0) = 0
            if (g_noisy) g_noisy[i * 3 + c] = s;
            if (g_gt && s != 0.f) atomicAdd(g_gt + j * 3 + c, -s);
        }
    }
}

int launch_reduce_partials(const float *partial, long long n, double scale, float *out, cudaStream_t st);   // warp_photo.cu

}  // namespace e2e

using namespace e2e;

extern "C" {

int e2e_knn1_fwd(const float *query, const float *transform, const float *ref, long long P1, long long P2,
                 float *dist2, long long *idx, void *stream)
{
    E2E_REQUIRE(query && ref && dist2 && idx && P1 > 0 && P2 > 0, "knn1: empty or null point cloud (P1=%lld, P2=%lld)", P1, P2);
    const long long per = (long long)KNN_NT * KNN_QPT;
    const long long blocks = (P1 + per - 1) / per;
    E2E_REQUIRE(blocks < (1ll << 31), "knn1: too many query points");
    knn1_fwd_kernel<<<(unsigned)blocks, KNN_NT, 0, (cudaStream_t)stream>>>(query, transform, ref, P1, P2, dist2, idx);
    count_launch();
    return finish_launch("knn1_fwd");
}

int e2e_knn1_bwd(const float *query, const float *transform, const float *ref, long long P1, long long P2,
                 const long long *idx, const float *grad_dist2, float *grad_query, float *grad_ref, void *stream)
{
    (void)P2;
    E2E_REQUIRE(query && ref && idx && grad_dist2 && (grad_query || grad_ref) && P1 > 0, "knn1_bwd: bad arguments");
    long long blocks = (P1 + KNN_NT - 1) / KNN_NT;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    knn1_bwd_kernel<<<(unsigned)blocks, KNN_NT, 0, (cudaStream_t)stream>>>(query, transform, ref, P1, idx, grad_dist2, grad_query, grad_ref);
    count_launch();
    return finish_launch("knn1_bwd");
}

static unsigned points_grid(long long n)
{
    long long blocks = (n + KNN_NT - 1) / KNN_NT;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

int e2e_transform_points_fwd(const float *points, const float *transform, long long n, float *out, void *stream)
{
    E2E_REQUIRE(transform && n >= 0 && (n == 0 || (points && out)), "transform_points: bad arguments");
    if (n == 0) return 0;
    transform_points_fwd_kernel<<<points_grid(n), KNN_NT, 0, (cudaStream_t)stream>>>(points, transform, n, out);
    count_launch();
    return finish_launch("transform_points_fwd");
}

int e2e_transform_points_bwd(const float *transform, const float *grad_out, long long n, float *grad_points, void *stream)
{
    E2E_REQUIRE(transform && n >= 0 && (n == 0 || (grad_out && grad_points)), "transform_points_bwd: bad arguments");
    if (n == 0) return 0;
    transform_points_bwd_kernel<<<points_grid(n), KNN_NT, 0, (cudaStream_t)stream>>>(transform, grad_out, n, grad_points);
    count_launch();
    return finish_launch("transform_points_bwd");
}

size_t e2e_color_points_workspace_bytes(long long P1) { return (size_t)points_grid(P1) * sizeof(float) + 256; }

int e2e_color_points_fwd(const float *gt_colors, const float *noisy_colors, const long long *idx, long long P1, float *loss,
                         void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(gt_colors && noisy_colors && idx && loss && P1 > 0, "color_points: bad arguments");
    const unsigned grid = points_grid(P1);
    E2E_REQUIRE(workspace && workspace_bytes >= grid * sizeof(float), "color_points: workspace too small");
    color_points_fwd_kernel<<<grid, KNN_NT, 0, (cudaStream_t)stream>>>(gt_colors, noisy_colors, idx, P1, (float *)workspace);
    count_launch();
    if (int rc = finish_launch("color_points_fwd")) return rc;
    return launch_reduce_partials((const float *)workspace, grid, 1.0 / (3.0 * (double)P1), loss, (cudaStream_t)stream);
}

int e2e_color_points_bwd(const float *gt_colors, const float *noisy_colors, const long long *idx, long long P1, const float *grad_loss,
                         float *grad_noisy, float *grad_gt, void *stream)
{
    E2E_REQUIRE(gt_colors && noisy_colors && idx && P1 > 0 && (grad_noisy || grad_gt), "color_points_bwd: bad arguments");
    color_points_bwd_kernel<<<points_grid(P1), KNN_NT, 0, (cudaStream_t)stream>>>(gt_colors, noisy_colors, idx, P1, grad_loss,
                                                                                  (float)(1.0 / (3.0 * (double)P1)), grad_noisy, grad_gt);
    count_launch();
    return finish_launch("color_points_bwd");
}

}  // extern "C"
