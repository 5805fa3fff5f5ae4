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

struct IcpIterRec {             // what one iteration leaves behind for the reverse sweep (e2e_icp_backward)
    double xi[6];               // Gauss-Newton / LM step of the iteration
    double lambda, q, e0, e1;   // damping the solve used, step gate, |b|^2 before / after the look-ahead
    double M[21], g[6];         // A^T A (upper triangle), A^T b
    float T[16];                // accumulated transform BEFORE the iteration's step
    float Sx[16];               // exp(xi): the look-ahead step (plain ICP: the committed step)
    float S[16];                // exp(q xi): the committed step
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

// sum over the block partials of column `lane`, in block order (deterministic), eight loads in flight at a time (a plain loop
// serialises ~75 L2 round trips in a one-warp kernel: 11 us per solve)
__device__ __forceinline__ double ordered_partial_sum(const float *partials, int nblk, int stride, int lane)
{
    double t = 0.0;
    for (int b = 0; b < nblk; b += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = (b + k < nblk) ? partials[(size_t)(b + k) * stride + lane] : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) t += (double)v[k];
    }
    return t;
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
                                                       float *errs, int it, IcpIterRec *rec)
{
    const int lane = threadIdx.x;
    if (lane < ICP_NV) st->sums[lane] = ordered_partial_sum(partials, nblk, ICP_NV, lane);      // fixed order: deterministic
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
        if (rec) {
            for (int k = 0; k < 6; k++) { rec->xi[k] = xi[k]; rec->g[k] = st->sums[21 + k]; }
            for (int k = 0; k < 21; k++) rec->M[k] = st->sums[k];
            rec->lambda = st->lambda; rec->q = 1.0; rec->e0 = st->sums[27]; rec->e1 = 0.0;
            for (int k = 0; k < 16; k++) { rec->T[k] = st->T[k]; rec->Sx[k] = st->step[k]; rec->S[k] = st->step[k]; }
        }
        if (!grad_icp) compose(st->step, st->T);
    } else {
        const double e0 = st->err0, e1 = st->sums[27];
        const double q = 1.0 / (1.0 + exp(-(e0 - e1) / nu));
        st->lambda = lambda_min + (lambda_max - lambda_min) / (1.0 + B * exp(-B2 * (e1 - e0) / nu));
        se3_exp_d(st->xi, q, st->step);
        if (rec) {
            rec->q = q; rec->e1 = e1;
            for (int k = 0; k < 16; k++) rec->S[k] = st->step[k];
        }
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

// History of a run for e2e_icp_backward: per iteration the record, the source cloud before the step, both sets of
// correspondences with their squared distances.
size_t e2e_icp_history_bytes(long long N, int numiters)
{
    if (N < 0) N = 0;
    if (numiters < 0) numiters = 0;
    const size_t K = (size_t)numiters;
    return a256(K * sizeof(IcpIterRec)) + a256(K * (size_t)N * 12) + 2 * a256(K * (size_t)N * 8) + 2 * a256(K * (size_t)N * 4) + 256;
}

}  // extern "C"

namespace e2e {

struct IcpHistory {
    IcpIterRec *rec;
    float *cur;                 // [K][N][3]
    long long *idx, *idx_t;     // [K][N]
    float *d2, *d2_t;           // [K][N]
};

static IcpHistory icp_history(void *history, long long N, int numiters)
{
    const size_t K = (size_t)numiters;
    unsigned char *w = (unsigned char *)history;
    IcpHistory h;
    h.rec = (IcpIterRec *)w;    w += a256(K * sizeof(IcpIterRec));
    h.cur = (float *)w;         w += a256(K * (size_t)N * 12);
    h.idx = (long long *)w;     w += a256(K * (size_t)N * 8);
    h.idx_t = (long long *)w;   w += a256(K * (size_t)N * 8);
    h.d2 = (float *)w;          w += a256(K * (size_t)N * 4);
    h.d2_t = (float *)w;
    return h;
}

static int icp_run(const float *src, long long N, const float *tgt, const float *tgt_normals, long long M,
                   const float *T_init, int numiters, float damp, float dist_thresh,
                   int grad_icp, float lambda_max, float B, float B2, float nu,
                   float *T_out, long long *idx_out, float *errs, void *workspace, size_t workspace_bytes, void *history, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    E2E_REQUIRE(src && tgt && tgt_normals && T_init && T_out && workspace, "icp: null argument");
    E2E_REQUIRE(N > 0 && M > 0 && numiters >= 0, "icp: empty point cloud or negative iteration count (N=%lld, M=%lld)", N, M);
    E2E_REQUIRE(workspace_bytes >= e2e_icp_workspace_bytes(N, M), "icp: workspace too small (e2e_icp_workspace_bytes)");
    E2E_REQUIRE(!grad_icp || nu != 0.0f, "icp: nu must be non-zero");
    unsigned char *w = (unsigned char *)workspace;
    void *grid = w;                     w += a256(e2e_knn1_grid_workspace_bytes(M));      // the target cloud, gridded once
    float *cur_ws = (float *)w;         w += a256((size_t)N * 12);
    float *trial = (float *)w;          w += a256((size_t)N * 12);
    float *dist2_ws = (float *)w;       w += a256((size_t)N * 4);
    long long *idx_ws = (long long *)w; w += a256((size_t)N * 8);
    float *dist2_t_ws = (float *)w;     w += a256((size_t)N * 4);       // look-ahead correspondences (GradICP)
    long long *idx_t_ws = (long long *)w;  w += a256((size_t)N * 8);
    float *partials = (float *)w;       w += a256((size_t)kNumSMs * 4 * ICP_NV * 4);
    IcpState *st = (IcpState *)w;
    IcpHistory h = {};
    if (history) h = icp_history(history, N, numiters);
    const int nb = icp_blocks(N);
    if (int rc = e2e_knn1_grid_build(tgt, M, grid, e2e_knn1_grid_workspace_bytes(M), stream)) return rc;
    icp_init_kernel<<<1, 32, 0, s>>>(st, T_init, (double)damp);
    float *cur = (history && numiters > 0) ? h.cur : cur_ws;
    icp_transform_kernel<<<nb, ICP_NT, 0, s>>>(src, T_init, cur, N);
    count_launch(2);
    long long *idx = idx_ws;
    for (int it = 0; it < numiters; it++) {
        // with a history every iteration keeps its own copy of the cloud / correspondences; otherwise they are overwritten
        idx = history ? h.idx + (size_t)it * N : idx_ws;
        float *dist2 = history ? h.d2 + (size_t)it * N : dist2_ws;
        long long *idx_t = history ? h.idx_t + (size_t)it * N : idx_t_ws;
        float *dist2_t = history ? h.d2_t + (size_t)it * N : dist2_t_ws;
        IcpIterRec *rec = history ? h.rec + it : nullptr;
        if (int rc = e2e_knn1_grid_query(cur, nullptr, N, M, dist2, idx, grid, stream)) return rc;
        icp_linearize_kernel<<<nb, ICP_NT, 0, s>>>(cur, tgt, tgt_normals, idx, dist2, dist_thresh, N, partials);
        icp_solve_kernel<<<1, 32, 0, s>>>(partials, nb, st, 0, grad_icp, damp, lambda_max, B, B2, nu, errs, it, rec);
        count_launch(2);
        if (grad_icp) {
            icp_transform_kernel<<<nb, ICP_NT, 0, s>>>(cur, st->step, trial, N);
            if (int rc = e2e_knn1_grid_query(trial, nullptr, N, M, dist2_t, idx_t, grid, stream)) return rc;
            icp_linearize_kernel<<<nb, ICP_NT, 0, s>>>(trial, tgt, tgt_normals, idx_t, dist2_t, dist_thresh, N, partials);
            icp_solve_kernel<<<1, 32, 0, s>>>(partials, nb, st, 1, grad_icp, damp, lambda_max, B, B2, nu, nullptr, it, rec);
            count_launch(3);
        }
        float *next = (history && it + 1 < numiters) ? h.cur + (size_t)(it + 1) * N * 3 : (history ? cur_ws : cur);
        icp_transform_kernel<<<nb, ICP_NT, 0, s>>>(cur, st->step, next, N);      // (in place without a history: one thread reads and writes its own point)
        cur = next;
        count_launch();
    }
    if (idx_out && numiters > 0) {      // correspondences of the last linearisation (what gradslam returns as chamfer_indices)
        if (cudaMemcpyAsync(idx_out, idx, (size_t)N * 8, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return finish_launch("icp: idx copy");
    }
    icp_finish_kernel<<<1, 32, 0, s>>>(st, T_out);
    count_launch();
    return finish_launch("icp_point_to_plane");
}

// =================================================================================================================
// Reverse mode of the iteration loop (GradICP's purpose: the recovered pose is differentiable w.r.t. the clouds; the
// correspondences are constants, as in gradslam).  Iteration k of the forward:
//     A_i = [n_i, s_i x n_i], b_i = n_i . (d_i - s_i)        s = cur_k, (d, n) = nearest target point / normal
//     xi = (A^T A + lambda_k I)^-1 A^T b                      e0 = |b|^2
//     GradICP: e1 = |b'|^2 of the look-ahead cloud exp(xi) cur_k;  q = sigmoid((e0 - e1) / nu);  lambda_{k+1} = gate(e1 - e0)
//     cur_{k+1} = S cur_k,  T_{k+1} = S T_k,  S = exp(q xi)    (plain ICP: q = 1)
// The sweep carries the adjoints of T_{k+1}, lambda_{k+1} and cur_{k+1} back to those of iteration k in four launches:
//   solveA   S-bar from T-bar and the cloud, through d exp to xi-bar and q-bar, the gates to e0-bar / e1-bar, T-bar_k = S^T T-bar
//   trial    (GradICP) per point: look-ahead residual -> its cloud, target / normal gradients, Sx-bar partials
//   solveB   xi-bar += d exp^T Sx-bar;  g-bar = (A^T A + lambda I)^-1 xi-bar;  lambda-bar_k = -g-bar . xi
//   lin      per point: A-bar, b-bar -> cur-bar_k (+ S_R^T cur-bar_{k+1} + the look-ahead part), target / normal gradients,
//            and the partials of sum cur-bar_k [cur_{k-1}; 1]^T that the next solveA needs (k = 0: the source cloud -> T_init-bar)
// d exp is taken with 6-wide dual numbers in float64 (one thread; 12 outputs).
// =================================================================================================================
struct D6 {
    double v, d[6];
};
__device__ __forceinline__ D6 d6_const(double c)
{
    D6 r;
    r.v = c;
    for (int i = 0; i < 6; i++) r.d[i] = 0.0;
    return r;
}
__device__ __forceinline__ D6 operator+(const D6 &a, const D6 &b)
{
    D6 r;
    r.v = a.v + b.v;
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] + b.d[i];
    return r;
}
__device__ __forceinline__ D6 operator-(const D6 &a, const D6 &b)
{
    D6 r;
    r.v = a.v - b.v;
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] - b.d[i];
    return r;
}
__device__ __forceinline__ D6 operator*(const D6 &a, const D6 &b)
{
    D6 r;
    r.v = a.v * b.v;
    for (int i = 0; i < 6; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
__device__ __forceinline__ D6 operator/(const D6 &a, const D6 &b)
{
    D6 r;
    r.v = a.v / b.v;
    for (int i = 0; i < 6; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
__device__ __forceinline__ D6 d6_fn(const D6 &a, double f, double df)      // f(a) with f'(a) = df
{
    D6 r;
    r.v = f;
    for (int i = 0; i < 6; i++) r.d[i] = df * a.d[i];
    return r;
}

// out (3 x 4, row-major) = top rows of exp(zeta) with d / d zeta, zeta = scale * xi (same formulas as se3_exp_d)
__device__ void se3_exp_dual(const double xi[6], double scale, D6 out[12])
{
    D6 v[3], w[3];
    for (int i = 0; i < 3; i++) {
        v[i] = d6_const(xi[i] * scale); v[i].d[i] = 1.0;
        w[i] = d6_const(xi[3 + i] * scale); w[i].d[3 + i] = 1.0;
    }
    const D6 z = d6_const(0.0), one = d6_const(1.0);
    const D6 K[9] = {z, z - w[2], w[1], w[2], z, z - w[0], z - w[1], w[0], z};
    D6 K2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    const D6 t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    D6 a, b, c, d;
    if (t2.v < 1e-12) {
        a = one; b = z; c = d6_const(0.5); d = z;
    } else {
        const double tv = sqrt(t2.v);
        const D6 t = d6_fn(t2, tv, 0.5 / tv);
        const D6 st = d6_fn(t, sin(tv), cos(tv)), ct = d6_fn(t, cos(tv), -sin(tv));
        a = st / t; b = (one - ct) / t2; c = b; d = (t - st) / (t2 * t);
    }
    for (int i = 0; i < 3; i++) {
        D6 tv = z;
        for (int j = 0; j < 3; j++) {
            const D6 I = d6_const(i == j ? 1.0 : 0.0);
            out[i * 4 + j] = I + a * K[i * 3 + j] + b * K2[i * 3 + j];
            tv = tv + (I + c * K[i * 3 + j] + d * K2[i * 3 + j]) * v[j];
        }
        out[i * 4 + 3] = tv;
    }
}

struct IcpBwdState {
    double Tbar[12];            // adjoint of the accumulated transform (rows 0..2)
    double lambar;              // adjoint of the damping the NEXT iteration starts with
    double xibar[6], gbar[6];
    double e0bar, e1bar;
};

__global__ void icp_bwd_init_kernel(IcpBwdState *bs, const float *grad_T)
{
    if (threadIdx.x < 12) bs->Tbar[threadIdx.x] = (double)grad_T[threadIdx.x];
    if (threadIdx.x == 0) bs->lambar = 0.0;
}

__global__ void __launch_bounds__(32) icp_bwd_solveA_kernel(const IcpIterRec *rec, IcpBwdState *bs, const float *sbar_partials, int nblk,
                                                            int grad_icp, double lambda_min, double lambda_max, double B, double B2, double nu)
{
    __shared__ double Sb[12];
    const int lane = threadIdx.x;
    if (lane < 12) Sb[lane] = ordered_partial_sum(sbar_partials, nblk, 12, lane);      // sum cur-bar_{k+1} [cur_k; 1]^T, fixed order
    __syncwarp();
    if (lane != 0) return;
    double Tk[16], Sm[16];
    for (int e = 0; e < 16; e++) { Tk[e] = (double)rec->T[e]; Sm[e] = (double)rec->S[e]; }
    for (int r = 0; r < 3; r++)         // S-bar += T-bar_{k+1} T_k^T
        for (int c = 0; c < 4; c++) {
            double t = 0.0;
            for (int j = 0; j < 4; j++) t += bs->Tbar[r * 4 + j] * Tk[c * 4 + j];
            Sb[r * 4 + c] += t;
        }
    double Tn[12];                      // T-bar_k = S^T T-bar_{k+1}
    for (int m = 0; m < 3; m++)
        for (int j = 0; j < 4; j++) {
            double t = 0.0;
            for (int i = 0; i < 3; i++) t += Sm[i * 4 + m] * bs->Tbar[i * 4 + j];
            Tn[m * 4 + j] = t;
        }
    for (int e = 0; e < 12; e++) bs->Tbar[e] = Tn[e];
    D6 out[12];
    se3_exp_dual(rec->xi, rec->q, out);
    double zbar[6];
    for (int i = 0; i < 6; i++) {
        double t = 0.0;
        for (int e = 0; e < 12; e++) t += Sb[e] * out[e].d[i];
        zbar[i] = t;
    }
    if (grad_icp) {
        const double q = rec->q;
        double qbar = 0.0;
        for (int i = 0; i < 6; i++) { qbar += zbar[i] * rec->xi[i]; bs->xibar[i] = q * zbar[i]; }
        const double ubar = qbar * q * (1.0 - q);
        double e0bar = ubar / nu, e1bar = -ubar / nu;
        const double ew = exp(-B2 * (rec->e1 - rec->e0) / nu);
        const double dl_dw = -(lambda_max - lambda_min) * B * ew / ((1.0 + B * ew) * (1.0 + B * ew));
        const double wbar = bs->lambar * dl_dw;
        e1bar += wbar * (-B2 / nu);
        e0bar += wbar * (B2 / nu);
        bs->e0bar = e0bar;
        bs->e1bar = e1bar;
    } else {
        for (int i = 0; i < 6; i++) bs->xibar[i] = zbar[i];
        bs->e0bar = 0.0;
        bs->e1bar = 0.0;
    }
}

// look-ahead residual of iteration k -> cloud, target, normals, Sx-bar partials
__global__ void __launch_bounds__(ICP_NT) icp_bwd_trial_kernel(const float *cur, const float *tgt, const float *nrm, const long long *idx_t,
                                                               const float *dist2_t, float thresh, long long N, const IcpIterRec *rec,
                                                               const IcpBwdState *bs, float *curbar_part, float *partials, float *g_tgt, float *g_nrm)
{
    __shared__ float red[ICP_NT / 32][12];
    float acc[12];
#pragma unroll
    for (int e = 0; e < 12; e++) acc[e] = 0.f;
    const float *Sx = rec->Sx;
    const float e1bar = (float)bs->e1bar;
    for (long long i = (long long)blockIdx.x * ICP_NT + threadIdx.x; i < N; i += (long long)gridDim.x * ICP_NT) {
        float cb[3] = {0.f, 0.f, 0.f};
        if (!(thresh >= 0.f && !(dist2_t[i] < thresh))) {
            const float s[3] = {cur[i * 3], cur[i * 3 + 1], cur[i * 3 + 2]};
            float tr[3];
#pragma unroll
            for (int r = 0; r < 3; r++) tr[r] = xadd(xadd(xadd(xmul(Sx[r * 4], s[0]), xmul(Sx[r * 4 + 1], s[1])), xmul(Sx[r * 4 + 2], s[2])), Sx[r * 4 + 3]);
            const long long j = idx_t[i];
            const float n[3] = {nrm[j * 3], nrm[j * 3 + 1], nrm[j * 3 + 2]};
            const float d[3] = {tgt[j * 3] - tr[0], tgt[j * 3 + 1] - tr[1], tgt[j * 3 + 2] - tr[2]};
            const float b1 = n[0] * d[0] + n[1] * d[1] + n[2] * d[2];
            const float bb = 2.0f * b1 * e1bar;
            float tb[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                tb[r] = -n[r] * bb;
                acc[r * 4] += tb[r] * s[0]; acc[r * 4 + 1] += tb[r] * s[1]; acc[r * 4 + 2] += tb[r] * s[2]; acc[r * 4 + 3] += tb[r];
            }
#pragma unroll
            for (int c = 0; c < 3; c++) cb[c] = Sx[c] * tb[0] + Sx[4 + c] * tb[1] + Sx[8 + c] * tb[2];
            if (g_tgt) { atomicAdd(g_tgt + j * 3, n[0] * bb); atomicAdd(g_tgt + j * 3 + 1, n[1] * bb); atomicAdd(g_tgt + j * 3 + 2, n[2] * bb); }
            if (g_nrm) { atomicAdd(g_nrm + j * 3, d[0] * bb); atomicAdd(g_nrm + j * 3 + 1, d[1] * bb); atomicAdd(g_nrm + j * 3 + 2, d[2] * bb); }
        }
        curbar_part[i * 3] = cb[0]; curbar_part[i * 3 + 1] = cb[1]; curbar_part[i * 3 + 2] = cb[2];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < 12; e++) {
        const float v = warp_sum(acc[e]);
        if (lane == 0) red[wid][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float t = 0.f;
        for (int w = 0; w < ICP_NT / 32; w++) t += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * 12 + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32) icp_bwd_solveB_kernel(const IcpIterRec *rec, IcpBwdState *bs, const float *sx_partials, int nblk, int grad_icp)
{
    __shared__ double Sb[12];
    const int lane = threadIdx.x;
    if (lane < 12) Sb[lane] = grad_icp ? ordered_partial_sum(sx_partials, nblk, 12, lane) : 0.0;
    __syncwarp();
    if (lane != 0) return;
    double xb[6];
    for (int i = 0; i < 6; i++) xb[i] = bs->xibar[i];
    if (grad_icp) {
        D6 out[12];
        se3_exp_dual(rec->xi, 1.0, out);
        for (int i = 0; i < 6; i++) {
            double t = 0.0;
            for (int e = 0; e < 12; e++) t += Sb[e] * out[e].d[i];
            xb[i] += t;
        }
    }
    // g-bar = (A^T A + lambda I)^-1 xi-bar (the matrix is symmetric: the forward's elimination)
    double Mx[6][7];
    int e = 0;
    for (int r = 0; r < 6; r++)
        for (int c = r; c < 6; c++) {
            Mx[r][c] = Mx[c][r] = rec->M[e++];
        }
    for (int r = 0; r < 6; r++) {
        Mx[r][r] += rec->lambda;
        Mx[r][6] = xb[r];
    }
    for (int k = 0; k < 6; k++) {
        int piv = k;
        for (int r = k + 1; r < 6; r++)
            if (fabs(Mx[r][k]) > fabs(Mx[piv][k])) piv = r;
        if (piv != k)
            for (int c = 0; c < 7; c++) { const double t = Mx[k][c]; Mx[k][c] = Mx[piv][c]; Mx[piv][c] = t; }
        const double d = Mx[k][k];
        if (d == 0.0) continue;
        for (int r = k + 1; r < 6; r++) {
            const double f = Mx[r][k] / d;
            for (int c = k; c < 7; c++) Mx[r][c] -= f * Mx[k][c];
        }
    }
    double gb[6];
    for (int k = 5; k >= 0; k--) {
        double sacc = Mx[k][6];
        for (int c = k + 1; c < 6; c++) sacc -= Mx[k][c] * gb[c];
        gb[k] = (Mx[k][k] != 0.0) ? sacc / Mx[k][k] : 0.0;
    }
    double lb = 0.0;
    for (int i = 0; i < 6; i++) { bs->gbar[i] = gb[i]; lb -= gb[i] * rec->xi[i]; }
    bs->lambar = lb;            // adjoint of the damping THIS iteration used = what the previous iteration's gate produced
}

// linearisation of iteration k -> cur-bar_k, target / normal gradients, partials of sum cur-bar_k [prev; 1]^T
// (prev = cur_{k-1}, or the source cloud for k = 0, where R_init^T cur-bar_0 is the source gradient)
__global__ void __launch_bounds__(ICP_NT) icp_bwd_lin_kernel(const float *cur, const float *prev, const float *tgt, const float *nrm, const long long *idx,
                                                             const float *dist2, float thresh, long long N, const IcpIterRec *rec, const IcpBwdState *bs,
                                                             const float *curbar_next, const float *curbar_part, float *curbar_out, float *partials,
                                                             float *g_tgt, float *g_nrm, const float *T_init, float *g_src)
{
    __shared__ float red[ICP_NT / 32][12];
    float acc[12];
#pragma unroll
    for (int e = 0; e < 12; e++) acc[e] = 0.f;
    const float *S = rec->S;
    float xi[6], gb[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { xi[k] = (float)rec->xi[k]; gb[k] = (float)bs->gbar[k]; }
    const float e0bar = (float)bs->e0bar;
    for (long long i = (long long)blockIdx.x * ICP_NT + threadIdx.x; i < N; i += (long long)gridDim.x * ICP_NT) {
        float cb[3] = {0.f, 0.f, 0.f};
        if (curbar_next) {
            const float c0 = curbar_next[i * 3], c1 = curbar_next[i * 3 + 1], c2 = curbar_next[i * 3 + 2];
#pragma unroll
            for (int c = 0; c < 3; c++) cb[c] = S[c] * c0 + S[4 + c] * c1 + S[8 + c] * c2;      // S_R^T cur-bar_{k+1}
        }
        if (curbar_part) { cb[0] += curbar_part[i * 3]; cb[1] += curbar_part[i * 3 + 1]; cb[2] += curbar_part[i * 3 + 2]; }
        if (!(thresh >= 0.f && !(dist2[i] < thresh))) {
            const long long j = idx[i];
            const float s[3] = {cur[i * 3], cur[i * 3 + 1], cur[i * 3 + 2]};
            const float n[3] = {nrm[j * 3], nrm[j * 3 + 1], nrm[j * 3 + 2]};
            const float d[3] = {tgt[j * 3] - s[0], tgt[j * 3 + 1] - s[1], tgt[j * 3 + 2] - s[2]};
            const float a[6] = {n[0], n[1], n[2], s[1] * n[2] - s[2] * n[1], s[2] * n[0] - s[0] * n[2], s[0] * n[1] - s[1] * n[0]};
            const float b = n[0] * d[0] + n[1] * d[1] + n[2] * d[2];
            float xa = 0.f, ga = 0.f;
#pragma unroll
            for (int k = 0; k < 6; k++) { xa += xi[k] * a[k]; ga += gb[k] * a[k]; }
            float ab[6];                 // A-bar = (M-bar + M-bar^T) a + b g-bar,  M-bar = -g-bar xi^T
#pragma unroll
            for (int k = 0; k < 6; k++) ab[k] = -(gb[k] * xa + xi[k] * ga) + b * gb[k];
            const float bbar = ga + 2.0f * b * e0bar;
            const float cbar[3] = {ab[3], ab[4], ab[5]};     // adjoint of s x n
            float sb[3] = {n[1] * cbar[2] - n[2] * cbar[1], n[2] * cbar[0] - n[0] * cbar[2], n[0] * cbar[1] - n[1] * cbar[0]};
            float nb[3] = {cbar[1] * s[2] - cbar[2] * s[1], cbar[2] * s[0] - cbar[0] * s[2], cbar[0] * s[1] - cbar[1] * s[0]};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                nb[k] += ab[k] + d[k] * bbar;
                sb[k] -= n[k] * bbar;
                cb[k] += sb[k];
            }
            if (g_tgt) { atomicAdd(g_tgt + j * 3, n[0] * bbar); atomicAdd(g_tgt + j * 3 + 1, n[1] * bbar); atomicAdd(g_tgt + j * 3 + 2, n[2] * bbar); }
            if (g_nrm) { atomicAdd(g_nrm + j * 3, nb[0]); atomicAdd(g_nrm + j * 3 + 1, nb[1]); atomicAdd(g_nrm + j * 3 + 2, nb[2]); }
        }
        if (curbar_out) { curbar_out[i * 3] = cb[0]; curbar_out[i * 3 + 1] = cb[1]; curbar_out[i * 3 + 2] = cb[2]; }
        const float p0 = prev[i * 3], p1 = prev[i * 3 + 1], p2 = prev[i * 3 + 2];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            acc[r * 4] += cb[r] * p0; acc[r * 4 + 1] += cb[r] * p1; acc[r * 4 + 2] += cb[r] * p2; acc[r * 4 + 3] += cb[r];
        }
        if (g_src) {
#pragma unroll
            for (int c = 0; c < 3; c++) g_src[i * 3 + c] = T_init[c] * cb[0] + T_init[4 + c] * cb[1] + T_init[8 + c] * cb[2];
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < 12; e++) {
        const float v = warp_sum(acc[e]);
        if (lane == 0) red[wid][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float t = 0.f;
        for (int w = 0; w < ICP_NT / 32; w++) t += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * 12 + threadIdx.x] = t;
    }
}

// T_init-bar = T-bar_0 (through the compositions) + sum cur-bar_0 [src; 1]^T
__global__ void __launch_bounds__(32) icp_bwd_finish_kernel(const IcpBwdState *bs, const float *partials, int nblk, float *g_T)
{
    const int lane = threadIdx.x;
    if (lane < 12) {
        g_T[lane] = (float)(bs->Tbar[lane] + ordered_partial_sum(partials, nblk, 12, lane));
    } else if (lane < 16) {
        g_T[lane] = 0.f;
    }
}

}  // namespace e2e

extern "C" {

int e2e_icp_point_to_plane(const float *src, long long N, const float *tgt, const float *tgt_normals, long long M,
                           const float *T_init, int numiters, float damp, float dist_thresh,
                           int grad_icp, float lambda_max, float B, float B2, float nu,
                           float *T_out, long long *idx_out, float *errs, void *workspace, size_t workspace_bytes, void *stream)
{
    return icp_run(src, N, tgt, tgt_normals, M, T_init, numiters, damp, dist_thresh, grad_icp, lambda_max, B, B2, nu, T_out, idx_out, errs,
                   workspace, workspace_bytes, nullptr, stream);
}

int e2e_icp_point_to_plane_saved(const float *src, long long N, const float *tgt, const float *tgt_normals, long long M,
                                 const float *T_init, int numiters, float damp, float dist_thresh,
                                 int grad_icp, float lambda_max, float B, float B2, float nu,
                                 float *T_out, long long *idx_out, void *workspace, size_t workspace_bytes,
                                 void *history, size_t history_bytes, void *stream)
{
    E2E_REQUIRE(history && history_bytes >= e2e_icp_history_bytes(N, numiters), "icp: history buffer too small (e2e_icp_history_bytes)");
    return icp_run(src, N, tgt, tgt_normals, M, T_init, numiters, damp, dist_thresh, grad_icp, lambda_max, B, B2, nu, T_out, idx_out, nullptr,
                   workspace, workspace_bytes, history, stream);
}

size_t e2e_icp_backward_workspace_bytes(long long N)
{
    if (N < 0) N = 0;
    return 3 * a256((size_t)N * 12) + 2 * a256((size_t)kNumSMs * 4 * 12 * 4) + a256(sizeof(IcpBwdState)) + 256;
}

int e2e_icp_backward(const float *src, long long N, const float *tgt, const float *tgt_normals, long long M,
                     const float *T_init, int numiters, float damp, float dist_thresh,
                     int grad_icp, float lambda_max, float B, float B2, float nu,
                     const void *history, const float *grad_T_out,
                     float *grad_src, float *grad_tgt, float *grad_normals, float *grad_T_init,
                     void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    E2E_REQUIRE(src && tgt && tgt_normals && T_init && history && grad_T_out && workspace, "icp_backward: null argument");
    E2E_REQUIRE(N > 0 && M > 0 && numiters >= 0, "icp_backward: empty point cloud or negative iteration count");
    E2E_REQUIRE(workspace_bytes >= e2e_icp_backward_workspace_bytes(N), "icp_backward: workspace too small");
    E2E_REQUIRE(!grad_icp || nu != 0.0f, "icp_backward: nu must be non-zero");
    const IcpHistory h = icp_history(const_cast<void *>(history), N, numiters);
    unsigned char *w = (unsigned char *)workspace;
    float *cb[2];
    cb[0] = (float *)w;                 w += a256((size_t)N * 12);
    cb[1] = (float *)w;                 w += a256((size_t)N * 12);
    float *cb_part = (float *)w;        w += a256((size_t)N * 12);
    float *sbar_partials = (float *)w;  w += a256((size_t)kNumSMs * 4 * 12 * 4);
    float *sx_partials = (float *)w;    w += a256((size_t)kNumSMs * 4 * 12 * 4);
    IcpBwdState *bs = (IcpBwdState *)w;
    const int nb = icp_blocks(N);
    // target / normal gradients are accumulated with atomics: the caller passes zeroed buffers (or NULL)
    icp_bwd_init_kernel<<<1, 32, 0, s>>>(bs, grad_T_out);
    count_launch();
    if (numiters == 0) {
        if (grad_src && cudaMemsetAsync(grad_src, 0, (size_t)N * 12, s) != cudaSuccess) return finish_launch("icp_backward: memset");
        if (grad_T_init) { icp_bwd_finish_kernel<<<1, 32, 0, s>>>(bs, sbar_partials, 0, grad_T_init); count_launch(); }
        return finish_launch("icp_backward");
    }
    int have_sbar = 0;                  // the last iteration's output cloud has no adjoint of its own
    for (int k = numiters - 1; k >= 0; k--) {
        const IcpIterRec *rec = h.rec + k;
        const float *cur = h.cur + (size_t)k * N * 3;
        const float *prev = k > 0 ? h.cur + (size_t)(k - 1) * N * 3 : src;
        float *cb_next = have_sbar ? cb[(k + 1) & 1] : nullptr, *cb_out = cb[k & 1];
        icp_bwd_solveA_kernel<<<1, 32, 0, s>>>(rec, bs, sbar_partials, have_sbar ? nb : 0, grad_icp, (double)damp, (double)lambda_max, (double)B, (double)B2, (double)nu);
        if (grad_icp)
            icp_bwd_trial_kernel<<<nb, ICP_NT, 0, s>>>(cur, tgt, tgt_normals, h.idx_t + (size_t)k * N, h.d2_t + (size_t)k * N, dist_thresh, N, rec, bs,
                                                       cb_part, sx_partials, grad_tgt, grad_normals);
        icp_bwd_solveB_kernel<<<1, 32, 0, s>>>(rec, bs, sx_partials, nb, grad_icp);
        icp_bwd_lin_kernel<<<nb, ICP_NT, 0, s>>>(cur, prev, tgt, tgt_normals, h.idx + (size_t)k * N, h.d2 + (size_t)k * N, dist_thresh, N, rec, bs,
                                                 cb_next, grad_icp ? cb_part : nullptr, cb_out, sbar_partials, grad_tgt, grad_normals,
                                                 k == 0 ? T_init : nullptr, k == 0 ? grad_src : nullptr);
        count_launch(grad_icp ? 4 : 3);
        have_sbar = 1;
    }
    if (grad_T_init) { icp_bwd_finish_kernel<<<1, 32, 0, s>>>(bs, sbar_partials, nb, grad_T_init); count_launch(); }
    return finish_launch("icp_backward");
}

}  // extern "C"
