// Point-to-plane ICP / GradICP odometry on sm_100a (gradslam's PointFusion with odom = "icp" / "gradicp": the reference's
// shipped default, configs/config.yaml:30-34, train_depth.py:111-116, online_adaption.py:362-363).
// Semantics: oracle/icp_oracle.py (restated from the gradSLAM paper; gradslam itself is not vendored by the reference).
//
// The whole iteration loop is enqueued by ONE call and never synchronises with the host: the current transform, the
// damping and the Gauss-Newton step live in device memory.  Per iteration
//     knn1 (csrc/knn_grid.cu: the target is gridded once per call, exact)  ->  linearize (Jacobian rows + 6x6 normal equations, block partials)
//     ->  solve (fixed-order fp64 reduction of the partials, 6x6 solve, se3 exponential, transform update)
//     ->  transform of the source cloud;
// GradICP adds the look-ahead (trial transform, knn1, linearize) and the logistic gates of step and damping.
#include "common.cuh"

namespace e2e {

constexpr int ICP_NT = 256;
constexpr int ICP_NV = 29;      // 21 upper-triangle entries of A^T A, 6 of A^T b, |b|^2, number of pairs

struct IcpState {               // device-resident loop state
    double sums[ICP_NV];
    double xi[6];
    double err0;
    double lambda;
    float T[16];                // accumulated transform
    float step[16];             // transform applied to the source cloud next
};

__global__ void __launch_bounds__(ICP_NT) icp_transform_kernel(const float *in, const float *T, float *out, long long N)
{
    for (long long i = (long long)blockIdx.x * ICP_NT + threadIdx.x; i < N; i += (long long)gridDim.x * ICP_NT) {
        const float x = in[i * 3], y = in[i * 3 + 1], z = in[i * 3 + 2];
#pragma unroll
        for (int r = 0; r < 3; r++)     // R p + t, left to right (transform_pointcloud of the oracle)
            out[i * 3 + r] = xadd(xadd(xadd(xmul(T[r * 4], x), xmul(T[r * 4 + 1], y)), xmul(T[r * 4 + 2], z)), T[r * 4 + 3]);
    }
}

__global__ void __launch_bounds__(ICP_NT) icp_linearize_kernel(const float *src, const float *tgt, const float *nrm, const long long *idx,
                                                               const float *dist2, float thresh, long long N, float *partials)
{
    __shared__ float red[ICP_NT / 32][ICP_NV];
    float acc[ICP_NV];
#pragma unroll
    for (int e = 0; e < ICP_NV; e++) acc[e] = 0.f;
    for (long long i = (long long)blockIdx.x * ICP_NT + threadIdx.x; i < N; i += (long long)gridDim.x * ICP_NT) {
        if (thresh >= 0.f && !(dist2[i] < thresh)) continue;
        const long long j = idx[i];
        const float sx = src[i * 3], sy = src[i * 3 + 1], sz = src[i * 3 + 2];
        const float nx = nrm[j * 3], ny = nrm[j * 3 + 1], nz = nrm[j * 3 + 2];
        const float dx = tgt[j * 3] - sx, dy = tgt[j * 3 + 1] - sy, dz = tgt[j * 3 + 2] - sz;
        float a[6] = {nx, ny, nz, sy * nz - sz * ny, sz * nx - sx * nz, sx * ny - sy * nx};
        const float b = nx * dx + ny * dy + nz * dz;
        int e = 0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = r; c < 6; c++) acc[e++] += a[r] * a[c];
#pragma unroll
        for (int r = 0; r < 6; r++) acc[21 + r] += a[r] * b;
        acc[27] += b * b;
        acc[28] += 1.0f;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < ICP_NV; e++) {
        const float v = warp_sum(acc[e]);
        if (lane == 0) red[wid][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < ICP_NV) {
        float t = 0.f;
        for (int w = 0; w < ICP_NT / 32; w++) t += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * ICP_NV + threadIdx.x] = t;
    }
}

__device__ void se3_exp_d(const double xi[6], double scale, float T[16])
{
    const double v[3] = {xi[0] * scale, xi[1] * scale, xi[2] * scale}, w[3] = {xi[3] * scale, xi[4] * scale, xi[5] * scale};
    const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double K2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double a, b, c, d;          // R = I + a K + b K^2, V = I + c K + d K^2
    if (t2 < 1e-12) {
        a = 1.0; b = 0.0; c = 0.5; d = 0.0;
    } else {
        const double t = sqrt(t2);
        a = sin(t) / t; b = (1.0 - cos(t)) / t2; c = b; d = (t - sin(t)) / (t2 * t);
    }
    for (int i = 0; i < 3; i++) {
        double tv = 0.0;
        for (int j = 0; j < 3; j++) {
            const double I = (i == j) ? 1.0 : 0.0;
            T[i * 4 + j] = (float)(I + a * K[i * 3 + j] + b * K2[i * 3 + j]);
            tv += (I + c * K[i * 3 + j] + d * K2[i * 3 + j]) * v[j];
        }
        T[i * 4 + 3] = (float)tv;
    }
    T[12] = T[13] = T[14] = 0.f;
    T[15] = 1.f;
}

__device__ void compose(const float step[16], float T[16])      // T <- step @ T
{
    float r[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += (double)step[i * 4 + k] * (double)T[k * 4 + j];
            r[i * 4 + j] = (float)s;
        }
    for (int e = 0; e < 16; e++) T[e] = r[e];
}

// phase 0: reduce, solve with the current damping; plain ICP commits the step, GradICP keeps it as the trial step.
// phase 1 (GradICP only): reduce the look-ahead residual, gate the step and the damping, commit.
__global__ void __launch_bounds__(32) icp_solve_kernel(const float *partials, int nblk, IcpState *st, int phase, int grad_icp,
                                                       double lambda_min, double lambda_max, double B, double B2, double nu,
                                                       float *errs, int it)
{
    const int lane = threadIdx.x;
    if (lane < ICP_NV) {        // fixed order: deterministic
        double t = 0.0;
        for (int b = 0; b < nblk; b++) t += (double)partials[(size_t)b * ICP_NV + lane];
        st->sums[lane] = t;
    }
    __syncwarp();
    if (lane != 0) return;
    if (phase == 0) {
        double M[6][7];
        int e = 0;
        for (int r = 0; r < 6; r++)
            for (int c = r; c < 6; c++) {
                M[r][c] = M[c][r] = st->sums[e++];
            }
        for (int r = 0; r < 6; r++) {
            M[r][r] += st->lambda;
            M[r][6] = st->sums[21 + r];
        }
        for (int k = 0; k < 6; k++) {       // Gaussian elimination with partial pivoting
            int piv = k;
            for (int r = k + 1; r < 6; r++)
                if (fabs(M[r][k]) > fabs(M[piv][k])) piv = r;
            if (piv != k)
                for (int c = 0; c < 7; c++) { const double t = M[k][c]; M[k][c] = M[piv][c]; M[piv][c] = t; }
            const double d = M[k][k];
            if (d == 0.0) continue;         // degenerate geometry: leave the component at 0
            for (int r = k + 1; r < 6; r++) {
                const double f = M[r][k] / d;
                for (int c = k; c < 7; c++) M[r][c] -= f * M[k][c];
            }
        }
        double xi[6];
        for (int k = 5; k >= 0; k--) {
            double s = M[k][6];
            for (int c = k + 1; c < 6; c++) s -= M[k][c] * xi[c];
            xi[k] = (M[k][k] != 0.0) ? s / M[k][k] : 0.0;
        }
        for (int k = 0; k < 6; k++) st->xi[k] = xi[k];
        st->err0 = st->sums[27];
        if (errs) errs[it] = (float)st->sums[27];
        se3_exp_d(xi, 1.0, st->step);
        if (!grad_icp) compose(st->step, st->T);
    } else {
        const double e0 = st->err0, e1 = st->sums[27];
        const double q = 1.0 / (1.0 + exp(-(e0 - e1) / nu));
        st->lambda = lambda_min + (lambda_max - lambda_min) / (1.0 + B * exp(-B2 * (e1 - e0) / nu));
        se3_exp_d(st->xi, q, st->step);
        compose(st->step, st->T);
    }
}

__global__ void icp_init_kernel(IcpState *st, const float *T_init, double lambda)
{
    if (threadIdx.x < 16) st->T[threadIdx.x] = T_init[threadIdx.x];
    if (threadIdx.x == 0) st->lambda = lambda;
}

__global__ void icp_finish_kernel(const IcpState *st, float *T_out)
{
    if (threadIdx.x < 16) T_out[threadIdx.x] = st->T[threadIdx.x];
}

static size_t a256(size_t n) { return (n + 255) / 256 * 256; }
static int icp_blocks(long long N)
{
    long long b = (N + ICP_NT - 1) / ICP_NT;
    if (b > kNumSMs * 4) b = kNumSMs * 4;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace e2e

using namespace e2e;

extern "C" {

size_t e2e_icp_workspace_bytes(long long N, long long M)
{
    if (N < 0) N = 0;
    return a256(e2e_knn1_grid_workspace_bytes(M)) + 2 * a256((size_t)N * 12) + 2 * a256((size_t)N * 4) + 2 * a256((size_t)N * 8) + a256((size_t)kNumSMs * 4 * ICP_NV * 4) +
           a256(sizeof(IcpState)) + 256;
}

int e2e_icp_point_to_plane(const float *src, long long N, const float *tgt, const float *tgt_normals, long long M,
                           const float *T_init, int numiters, float damp, float dist_thresh,
                           int grad_icp, float lambda_max, float B, float B2, float nu,
                           float *T_out, long long *idx_out, float *errs, void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    E2E_REQUIRE(src && tgt && tgt_normals && T_init && T_out && workspace, "icp: null argument");
    E2E_REQUIRE(N > 0 && M > 0 && numiters >= 0, "icp: empty point cloud or negative iteration count (N=%lld, M=%lld)", N, M);
    E2E_REQUIRE(workspace_bytes >= e2e_icp_workspace_bytes(N, M), "icp: workspace too small (e2e_icp_workspace_bytes)");
    E2E_REQUIRE(!grad_icp || nu != 0.0f, "icp: nu must be non-zero");
    unsigned char *w = (unsigned char *)workspace;
    void *grid = w;                     w += a256(e2e_knn1_grid_workspace_bytes(M));      // the target cloud, gridded once
    float *cur = (float *)w;            w += a256((size_t)N * 12);
    float *trial = (float *)w;          w += a256((size_t)N * 12);
    float *dist2 = (float *)w;          w += a256((size_t)N * 4);
    long long *idx = (long long *)w;    w += a256((size_t)N * 8);
    float *dist2_t = (float *)w;        w += a256((size_t)N * 4);       // look-ahead correspondences (GradICP)
    long long *idx_t = (long long *)w;  w += a256((size_t)N * 8);
    float *partials = (float *)w;       w += a256((size_t)kNumSMs * 4 * ICP_NV * 4);
    IcpState *st = (IcpState *)w;
    const int nb = icp_blocks(N);
    if (int rc = e2e_knn1_grid_build(tgt, M, grid, e2e_knn1_grid_workspace_bytes(M), stream)) return rc;
    icp_init_kernel<<<1, 32, 0, s>>>(st, T_init, (double)damp);
    icp_transform_kernel<<<nb, ICP_NT, 0, s>>>(src, T_init, cur, N);
    count_launch(2);
    for (int it = 0; it < numiters; it++) {
        if (int rc = e2e_knn1_grid_query(cur, nullptr, N, M, dist2, idx, grid, stream)) return rc;
        icp_linearize_kernel<<<nb, ICP_NT, 0, s>>>(cur, tgt, tgt_normals, idx, dist2, dist_thresh, N, partials);
        icp_solve_kernel<<<1, 32, 0, s>>>(partials, nb, st, 0, grad_icp, damp, lambda_max, B, B2, nu, errs, it);
        count_launch(2);
        if (grad_icp) {
            icp_transform_kernel<<<nb, ICP_NT, 0, s>>>(cur, st->step, trial, N);
            if (int rc = e2e_knn1_grid_query(trial, nullptr, N, M, dist2_t, idx_t, grid, stream)) return rc;
            icp_linearize_kernel<<<nb, ICP_NT, 0, s>>>(trial, tgt, tgt_normals, idx_t, dist2_t, dist_thresh, N, partials);
            icp_solve_kernel<<<1, 32, 0, s>>>(partials, nb, st, 1, grad_icp, damp, lambda_max, B, B2, nu, nullptr, it);
            count_launch(3);
        }
        icp_transform_kernel<<<nb, ICP_NT, 0, s>>>(cur, st->step, cur, N);      // in place: one thread reads and writes its own point
        count_launch();
    }
    if (idx_out && numiters > 0) {      // correspondences of the last linearisation (what gradslam returns as chamfer_indices)
        if (cudaMemcpyAsync(idx_out, idx, (size_t)N * 8, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return finish_launch("icp: idx copy");
    }
    icp_finish_kernel<<<1, 32, 0, s>>>(st, T_out);
    count_launch();
    return finish_launch("icp_point_to_plane");
}

}  // extern "C"
