// PointFusion on sm_100a: per-frame vertex / normal / confidence maps, projective data association as a
// per-pixel 64-bit atomicMin "index map", confidence-weighted merge and order-preserving append.
//
// Semantics: gradslam's RGBDImages + slam/fusionutils.py as restated in oracle/fusion_oracle.py (SURVEY.md
// appendix B; gradslam itself is not vendored by the reference).  All geometry that feeds an integer
// decision (frustum test, rounding to a pixel, distance / normal thresholds, best-candidate ordering) is
// computed with one IEEE rounding per operation in the oracle's written order (__fmul_rn / __fadd_rn, no
// FMA contraction), which is what makes the index map and the append order bit-exact.
//
// The number of map points lives on the device (`n_map`, int64): a step never synchronises with the host.
// The host only supplies an upper bound `n_upper` used to size grids; threads beyond *n_map exit.
#include "common.cuh"

namespace e2e {

constexpr int FU_NT = 256;
constexpr int SCAN_CHUNK = 256;        // pixels per CTA in the append scan (row-major chunks): one pixel per thread

struct Cam {
    float ifx, ify, icx, icy;          // closed-form inverse intrinsics
    float R[9], t[3];                  // pose (camera -> world)
    float Ri[9], ti[3];                // inverse pose
    float K[12];                       // rows 0..2 of the intrinsics (general 3x4)
};

// One thread builds the per-frame constants from K (4x4) and pose (4x4).
__device__ __forceinline__ void build_cam(const float *K, const float *pose, Cam &c)
{
    const float fx = K[0], fy = K[5], cx = K[2], cy = K[6];
    c.ifx = xdiv(1.0f, fx);
    c.ify = xdiv(1.0f, fy);
    c.icx = -xdiv(cx, fx);
    c.icy = -xdiv(cy, fy);
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            c.R[i * 3 + j] = pose[i * 4 + j];
            c.Ri[j * 3 + i] = pose[i * 4 + j];
        }
        c.t[i] = pose[i * 4 + 3];
#pragma unroll
        for (int j = 0; j < 4; j++) c.K[i * 4 + j] = K[i * 4 + j];
    }
#pragma unroll
    for (int i = 0; i < 3; i++)        // -R^T t, accumulated left to right
        c.ti[i] = -xadd(xadd(xmul(c.Ri[i * 3 + 0], c.t[0]), xmul(c.Ri[i * 3 + 1], c.t[1])), xmul(c.Ri[i * 3 + 2], c.t[2]));
}

__device__ __forceinline__ void stage_cam(const float *K, const float *pose, Cam *sc)
{
    if (threadIdx.x == 0) build_cam(K, pose, *sc);
    __syncthreads();
}

__device__ __forceinline__ float dot3_lr(float a0, float a1, float a2, float b0, float b1, float b2)
{
    return xadd(xadd(xmul(a0, b0), xmul(a1, b1)), xmul(a2, b2));
}

// masked local vertex of pixel (y, x) with depth d:  ((u*ifx + icx) * d, (v*ify + icy) * d, d) * [d > 0]
__device__ __forceinline__ void local_vertex(float d, int y, int x, const Cam &c, float V[3], float &m)
{
    m = (d > 0.0f) ? 1.0f : 0.0f;
    V[0] = xmul(xmul(xadd(xmul((float)x, c.ifx), c.icx), d), m);
    V[1] = xmul(xmul(xadd(xmul((float)y, c.ify), c.icy), d), m);
    V[2] = xmul(d, m);
}

// One pixel of the per-frame maps from its depth d and the depths of its right (dr) and lower (dd) neighbours
// (forward differences; the last column / row has none).  Shared by rgbd_maps_kernel and the whole-sequence kernel.
struct PixelMaps {
    float vg[3], ng[3], alpha, valid;
};

__device__ __forceinline__ void rgbd_pixel_core(float d, float dr, float dd, int H, int W, int y, int x, const Cam &c,
                                                float two_sigma2, PixelMaps &o)
{
    float V[3], Vr[3], Vd[3], m, mr, md;
    local_vertex(d, y, x, c, V, m);
    float dh[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f};
    if (x + 1 < W) {
        local_vertex(dr, y, x + 1, c, Vr, mr);
#pragma unroll
        for (int k = 0; k < 3; k++) dh[k] = xsub(Vr[k], V[k]);
    }
    if (y + 1 < H) {
        local_vertex(dd, y + 1, x, c, Vd, md);
#pragma unroll
        for (int k = 0; k < 3; k++) dv[k] = xsub(Vd[k], V[k]);
    }
    float n[3];
    n[0] = xsub(xmul(dh[1], dv[2]), xmul(dh[2], dv[1]));
    n[1] = xsub(xmul(dh[2], dv[0]), xmul(dh[0], dv[2]));
    n[2] = xsub(xmul(dh[0], dv[1]), xmul(dh[1], dv[0]));
    float norm = __fsqrt_rn(xadd(xadd(xmul(n[0], n[0]), xmul(n[1], n[1])), xmul(n[2], n[2])));
    if (norm == 0.0f) norm = 1.0f;
    float N[3];
#pragma unroll
    for (int k = 0; k < 3; k++) N[k] = xmul(xdiv(n[k], norm), m);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float vg = xadd(dot3_lr(c.R[k * 3], c.R[k * 3 + 1], c.R[k * 3 + 2], V[0], V[1], V[2]), c.t[k]);
        o.vg[k] = xmul(vg, m);
        o.ng[k] = dot3_lr(c.R[k * 3], c.R[k * 3 + 1], c.R[k * 3 + 2], N[0], N[1], N[2]);
    }
    // alpha = exp(-(X^2 + Y^2) / (2 sigma^2)): the float32 argument is exponentiated in double and
    // rounded once, so host (numpy) and device agree on the bits of the confidence counts.
    const float arg = -xdiv(xadd(xmul(V[0], V[0]), xmul(V[1], V[1])), two_sigma2);
    o.alpha = (float)exp((double)arg);
    o.valid = m;
}

__device__ __forceinline__ void rgbd_pixel(const float *depth, int H, int W, int i, const Cam &c, float two_sigma2,
                                           float *vertex_g, float *normal_g, float *alpha, unsigned char *valid)
{
    const int y = i / W, x = i - y * W;
    const float d = depth[i], dr = (x + 1 < W) ? depth[i + 1] : 0.0f, dd = (y + 1 < H) ? depth[i + W] : 0.0f;
    PixelMaps o;
    rgbd_pixel_core(d, dr, dd, H, W, y, x, c, two_sigma2, o);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        vertex_g[i * 3 + k] = o.vg[k];
        normal_g[i * 3 + k] = o.ng[k];
    }
    alpha[i] = o.alpha;
    valid[i] = (o.valid != 0.0f) ? 1 : 0;
}

__global__ void __launch_bounds__(FU_NT) rgbd_maps_kernel(const float *depth, const float *K, const float *pose, int H, int W,
                                                          float two_sigma2, float *vertex_g, float *normal_g, float *alpha,
                                                          unsigned char *valid)
{
    __shared__ Cam c;
    stage_cam(K, pose, &c);
    const int HW = H * W;
    for (int i = blockIdx.x * FU_NT + threadIdx.x; i < HW; i += gridDim.x * FU_NT)
        rgbd_pixel(depth, H, W, i, c, two_sigma2, vertex_g, normal_g, alpha, valid);
}

// d(vertex_g, alpha)/d depth.  Normals carry no gradient (no loss in the reference reads them).
__global__ void __launch_bounds__(FU_NT) rgbd_maps_bwd_kernel(const float *depth, const float *K, const float *pose, int H, int W,
                                                              float two_sigma2, const float *g_vertex_g, const float *g_alpha,
                                                              float *g_depth)
{
    __shared__ Cam c;
    stage_cam(K, pose, &c);
    const int HW = H * W;
    for (int i = blockIdx.x * FU_NT + threadIdx.x; i < HW; i += gridDim.x * FU_NT) {
        const int y = i / W, x = i - y * W;
        const float d = depth[i];
        float g = 0.f;
        if (d > 0.0f) {
            const float rx = (float)x * c.ifx + c.icx, ry = (float)y * c.ify + c.icy;
            if (g_vertex_g) {
                const float g0 = g_vertex_g[i * 3], g1 = g_vertex_g[i * 3 + 1], g2 = g_vertex_g[i * 3 + 2];
                // local gradient = R^T g
                const float l0 = c.R[0] * g0 + c.R[3] * g1 + c.R[6] * g2;
                const float l1 = c.R[1] * g0 + c.R[4] * g1 + c.R[7] * g2;
                const float l2 = c.R[2] * g0 + c.R[5] * g1 + c.R[8] * g2;
                g += l0 * rx + l1 * ry + l2;
            }
            if (g_alpha) {
                const float X = rx * d, Y = ry * d;
                const float a = (float)exp((double)(-(X * X + Y * Y) / two_sigma2));
                g += g_alpha[i] * a * (-2.0f * (X * rx + Y * ry) / two_sigma2);
            }
        }
        g_depth[i] = g;
    }
}

// ---------------------------------------------------------------------------------------------
// e2e_fusion_associate
// ---------------------------------------------------------------------------------------------
struct AssocParams {
    const float *pts, *nrm, *cc;
    const long long *n_map;
    long long n_upper;
    const float *K, *pose, *vertex_g, *normal_g;
    int H, W;
    float dist_th, dot_th, u_hi, v_hi;
    unsigned long long *keys, *index_map;
    int *cand;
};

// Steps 1-3 of the association for one map point, split at the two memory round trips.
//   assoc_pixel    active map points: into the live camera, in front, inside the frustum, rounded to a pixel (-1 = inactive)
//   assoc_key      similar (close in space, similar normal) -> the 64-bit key (1/(c + 1e-20), dist^2) it competes with
struct AssocConst {
    int H, W;
    float dist_th, dot_th, u_hi, v_hi;
};

__device__ __forceinline__ int assoc_pixel(const Cam &c, const AssocConst &a, float px, float py, float pz)
{
    const float qx = xadd(dot3_lr(c.Ri[0], c.Ri[1], c.Ri[2], px, py, pz), c.ti[0]);
    const float qy = xadd(dot3_lr(c.Ri[3], c.Ri[4], c.Ri[5], px, py, pz), c.ti[1]);
    const float qz = xadd(dot3_lr(c.Ri[6], c.Ri[7], c.Ri[8], px, py, pz), c.ti[2]);
    if (!(qz > 0.0f)) return -1;
    const float h0 = xadd(dot3_lr(c.K[0], c.K[1], c.K[2], qx, qy, qz), c.K[3]);
    const float h1 = xadd(dot3_lr(c.K[4], c.K[5], c.K[6], qx, qy, qz), c.K[7]);
    const float h2 = xadd(dot3_lr(c.K[8], c.K[9], c.K[10], qx, qy, qz), c.K[11]);
    const float u = xdiv(h0, h2), v = xdiv(h1, h2);
    if (!(u > -1e-3f && u < a.u_hi && v > -1e-3f && v < a.v_hi)) return -1;
    int w = __float2int_rn(u), h = __float2int_rn(v);       // round half to even, like torch.round
    w = min(max(w, 0), a.W - 1);
    h = min(max(h, 0), a.H - 1);
    return h * a.W + w;
}

__device__ __forceinline__ unsigned long long assoc_make_key(float cc, float dist2)
{
    const float inv_c = xdiv(1.0f, xadd(cc, 1e-20f));
    return ((unsigned long long)__float_as_uint(inv_c) << 32) | (unsigned long long)__float_as_uint(dist2);
}

__device__ __forceinline__ float assoc_dist2(float vx, float vy, float vz, float px, float py, float pz)
{
    const float dx = xsub(vx, px), dy = xsub(vy, py), dz = xsub(vz, pz);
    return xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
}

__device__ __forceinline__ bool assoc_key(const AssocConst &a, float px, float py, float pz, float nx, float ny, float nz, float cc,
                                          float vx, float vy, float vz, float gx, float gy, float gz, unsigned long long &key)
{
    const float dist2 = assoc_dist2(vx, vy, vz, px, py, pz);
    const float dot = dot3_lr(gx, gy, gz, nx, ny, nz);
    if (!(__fsqrt_rn(dist2) < a.dist_th && dot > a.dot_th)) return false;
    key = assoc_make_key(cc, dist2);        // best unique: lexicographic minimum of (1/(c + 1e-20), dist^2, n) per pixel
    return true;
}

// Pass 1: every map point is projected into the live frame; candidates (in the frustum, close in space, similar
// normal) race for their pixel with a 64-bit atomicMin of the key (1/(c+1e-20), dist^2) and remember their pixel in
// cand[n] (-1 = not a candidate).  All per-point loads (point, normal, confidence) are issued up front, so the
// dependent chain is two memory round trips (point data -> live-frame gather) instead of four.
__global__ void __launch_bounds__(FU_NT) associate_pass1_kernel(const AssocParams p)
{
    __shared__ Cam c;
    stage_cam(p.K, p.pose, &c);
    const AssocConst a{p.H, p.W, p.dist_th, p.dot_th, p.u_hi, p.v_hi};
    const long long N = *p.n_map;
    for (long long n = (long long)blockIdx.x * FU_NT + threadIdx.x; n < N; n += (long long)gridDim.x * FU_NT) {
        const float px = p.pts[n * 3], py = p.pts[n * 3 + 1], pz = p.pts[n * 3 + 2];
        const float nx = p.nrm[n * 3], ny = p.nrm[n * 3 + 1], nz = p.nrm[n * 3 + 2];
        const float cc = p.cc[n];
        int cand = -1;
        const int pix = assoc_pixel(c, a, px, py, pz);
        if (pix >= 0) {
            const float vx = p.vertex_g[pix * 3], vy = p.vertex_g[pix * 3 + 1], vz = p.vertex_g[pix * 3 + 2];
            const float gx = p.normal_g[pix * 3], gy = p.normal_g[pix * 3 + 1], gz = p.normal_g[pix * 3 + 2];
            unsigned long long key;
            if (assoc_key(a, px, py, pz, nx, ny, nz, cc, vx, vy, vz, gx, gy, gz, key)) {
                atomicMin(p.keys + pix, key);
                cand = pix;
            }
        }
        p.cand[n] = cand;
    }
}

// Pass 2: among the candidates whose key won their pixel, the smallest map index wins (ties on the full key).
// Non-candidates leave after one 4-byte load; candidates recompute their key with the same operations.
__global__ void __launch_bounds__(FU_NT) associate_pass2_kernel(const AssocParams p)
{
    const long long N = *p.n_map;
    for (long long n = (long long)blockIdx.x * FU_NT + threadIdx.x; n < N; n += (long long)gridDim.x * FU_NT) {
        const int pix = p.cand[n];
        if (pix < 0) continue;
        const float px = p.pts[n * 3], py = p.pts[n * 3 + 1], pz = p.pts[n * 3 + 2];
        const float dist2 = assoc_dist2(p.vertex_g[pix * 3], p.vertex_g[pix * 3 + 1], p.vertex_g[pix * 3 + 2], px, py, pz);
        const unsigned long long key = assoc_make_key(p.cc[n], dist2);
        if (p.keys[pix] == key) atomicMin(p.index_map + pix, (unsigned long long)n);
    }
}

// ---------------------------------------------------------------------------------------------
// e2e_fusion_active_points: gradslam.slam.fusionutils.find_active_map_points (imported at online_adaption.py:35; SURVEY 8(a) a18,
// appendix B step 1) as a stand-alone entry: rows (b, n, h, w) int64 of the map points that lie in front of the live camera and
// project into the frame, ORDERED BY n (stream compaction: count per chunk | scan | write).  Same arithmetic as the association
// kernels (assoc_pixel), so the rows are exactly the candidates those kernels consider.
// ---------------------------------------------------------------------------------------------
struct ActiveParams {
    const float *pts;
    long long n;
    const float *K, *pose;
    int H, W;
    float u_hi, v_hi;
    long long batch;
    int *pix;              // [n] pixel of every map point (-1 = inactive)
    int *chunk_counts;     // [nchunks], then exclusive offsets in place
    long long *rows;       // [n_active, 4]
};

__global__ void __launch_bounds__(FU_NT) active_count_kernel(const ActiveParams p)
{
    __shared__ Cam c;
    __shared__ int wsum[FU_NT / 32];
    stage_cam(p.K, p.pose, &c);
    const AssocConst a{p.H, p.W, 0.f, 0.f, p.u_hi, p.v_hi};
    const long long n = (long long)blockIdx.x * FU_NT + threadIdx.x;
    int pix = -1;
    if (n < p.n) {
        pix = assoc_pixel(c, a, p.pts[n * 3], p.pts[n * 3 + 1], p.pts[n * 3 + 2]);
        p.pix[n] = pix;
    }
    const int cnt = __popc(__ballot_sync(0xffffffffu, pix >= 0));
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < FU_NT / 32; w++) t += wsum[w];
        p.chunk_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(FU_NT) active_write_kernel(const ActiveParams p)
{
    __shared__ int woff[FU_NT / 32];
    const long long n = (long long)blockIdx.x * FU_NT + threadIdx.x;
    const int pix = (n < p.n) ? p.pix[n] : -1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned ballot = __ballot_sync(0xffffffffu, pix >= 0);
    if (lane == 0) woff[wid] = __popc(ballot);
    __syncthreads();
    if (pix < 0) return;
    long long slot = p.chunk_counts[blockIdx.x] + __popc(ballot & ((1u << lane) - 1));
    for (int w = 0; w < wid; w++) slot += woff[w];
    long long *r = p.rows + slot * 4;
    r[0] = p.batch;
    r[1] = n;
    r[2] = pix / p.W;
    r[3] = pix - (pix / p.W) * p.W;
}

// ---------------------------------------------------------------------------------------------
// e2e_fusion_merge_append: merge + count | scan of chunk counts | append
// ---------------------------------------------------------------------------------------------
struct FuseParams {
    float *pts, *nrm, *col, *cc;
    const long long *n_map;
    long long capacity;
    const float *vertex_g, *normal_g, *rgb, *alpha;
    const unsigned char *valid;
    const long long *index_map;
    int H, W;
    long long *append_slot, *n_out;
    int *chunk_counts;        // [nchunks], then exclusive offsets in place
    int nchunks;
};

__device__ __forceinline__ float merge_val(float c, float o, float a, float f, float den)
{
    return xdiv(xadd(xmul(c, o), xmul(a, f)), den);
}

// (c*old + a*new) / (c + a) for point, normal and colour of map entry n from live pixel i, one rounding per operation.
// All twenty loads are issued before the first store (the arrays may alias as far as the compiler knows, so
// interleaving loads and stores would serialise nine memory round trips).
__device__ __forceinline__ void merge_point(float *pts, float *nrm, float *col, float *ccount, long long n,
                                            const float *vertex_g, const float *normal_g, const float *rgb, const float *alpha, int i)
{
    const float c = ccount[n], a = alpha[i];
    float o[9], f[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        o[j] = pts[n * 3 + j]; o[3 + j] = nrm[n * 3 + j]; o[6 + j] = col[n * 3 + j];
        f[j] = vertex_g[i * 3 + j]; f[3 + j] = normal_g[i * 3 + j]; f[6 + j] = rgb[i * 3 + j];
    }
    const float den = xadd(c, a);
    float r[9];
#pragma unroll
    for (int j = 0; j < 9; j++) r[j] = merge_val(c, o[j], a, f[j], den);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        pts[n * 3 + j] = r[j]; nrm[n * 3 + j] = r[3 + j]; col[n * 3 + j] = r[6 + j];
    }
    ccount[n] = den;
}

__global__ void __launch_bounds__(FU_NT) fuse_merge_count_kernel(const FuseParams p)
{
    __shared__ int wsum[FU_NT / 32];
    const int HW = p.H * p.W;
    const int base = blockIdx.x * SCAN_CHUNK;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK / FU_NT; k++) {
        const int i = base + k * FU_NT + threadIdx.x;
        if (i >= HW) continue;
        const long long n = p.index_map[i];
        if (n >= 0) {
            merge_point(p.pts, p.nrm, p.col, p.cc, n, p.vertex_g, p.normal_g, p.rgb, p.alpha, i);
        } else if (p.valid[i]) {
            cnt++;
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < FU_NT / 32; w++) t += wsum[w];
        p.chunk_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) fuse_scan_kernel(int *chunk_counts, int nchunks, const long long *n_map, long long *n_out)
{
    __shared__ int wtot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < nchunks) ? chunk_counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wtot[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += t;
            }
            wtot[threadIdx.x] = wi - w;          // exclusive warp offsets
        }
        __syncthreads();
        const int excl = carry + wtot[threadIdx.x >> 5] + incl - v;
        if (i < nchunks) chunk_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_out[0] = (n_map ? n_map[0] : 0ll) + (long long)carry;
}

__global__ void __launch_bounds__(FU_NT) fuse_append_kernel(const FuseParams p)
{
    __shared__ int woff[FU_NT / 32];
    __shared__ int running;
    const int HW = p.H * p.W;
    const int base = blockIdx.x * SCAN_CHUNK;
    const long long slot0 = p.n_map[0] + (long long)p.chunk_counts[blockIdx.x];
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < SCAN_CHUNK / FU_NT; k++) {
        const int i = base + k * FU_NT + threadIdx.x;
        const bool flag = (i < HW) && (p.index_map[i] < 0) && p.valid[i];
        const unsigned ballot = __ballot_sync(0xffffffffu, flag);
        const int rank = __popc(ballot & ((1u << lane) - 1));
        if (lane == 0) woff[wid] = __popc(ballot);
        __syncthreads();
        int before = running;
        for (int w = 0; w < wid; w++) before += woff[w];
        if (i < HW) {
            long long slot = -1;
            if (flag) {
                slot = slot0 + before + rank;
                if (slot < p.capacity) {
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        p.pts[slot * 3 + j] = p.vertex_g[i * 3 + j];
                        p.nrm[slot * 3 + j] = p.normal_g[i * 3 + j];
                        p.col[slot * 3 + j] = p.rgb[i * 3 + j];
                    }
                    p.cc[slot] = p.alpha[i];
                }
            }
            if (p.append_slot) p.append_slot[i] = slot;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < FU_NT / 32; w++) t += woff[w];
            running += t;
        }
        __syncthreads();
    }
}

// Backward of merge + append w.r.t. the live frame (vertex_g, rgb, alpha) and the old map (points, colors,
// ccount).  Old-map gradients for unmatched points are a pass-through the caller has already copied into
// g_old_*; this kernel overwrites the matched entries.  Normals are not differentiated.
struct FuseBwdParams {
    const float *g_pts, *g_col, *g_cc;             // gradients w.r.t. the NEW map
    const float *old_pts, *old_col, *old_cc;       // the map that entered the step
    const float *vertex_g, *rgb, *alpha;
    const long long *index_map, *append_slot;
    int H, W;
    float *g_vertex_g, *g_rgb, *g_alpha;           // [H,W,3], [H,W,3], [H,W]
    float *g_old_pts, *g_old_col, *g_old_cc;       // nullable
};

__global__ void __launch_bounds__(FU_NT) fuse_bwd_kernel(const FuseBwdParams p)
{
    const int HW = p.H * p.W;
    for (int i = blockIdx.x * FU_NT + threadIdx.x; i < HW; i += gridDim.x * FU_NT) {
        float gv[3] = {0.f, 0.f, 0.f}, gr[3] = {0.f, 0.f, 0.f}, ga = 0.f;
        const long long s = p.append_slot[i], n = p.index_map[i];
        if (s >= 0) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                gv[j] = p.g_pts ? p.g_pts[s * 3 + j] : 0.f;
                gr[j] = p.g_col ? p.g_col[s * 3 + j] : 0.f;
            }
            ga = p.g_cc ? p.g_cc[s] : 0.f;
        } else if (n >= 0) {
            const float c = p.old_cc[n], a = p.alpha[i];
            const float rden = 1.0f / (c + a);
            const float wa = a * rden, wc = c * rden;
            float gc = p.g_cc ? p.g_cc[n] : 0.f;
            ga = gc;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const float op = p.old_pts[n * 3 + j], fp = p.vertex_g[i * 3 + j];
                const float oc = p.old_col[n * 3 + j], fc = p.rgb[i * 3 + j];
                const float np_ = wc * op + wa * fp, nc = wc * oc + wa * fc;
                const float g1 = p.g_pts ? p.g_pts[n * 3 + j] : 0.f, g2 = p.g_col ? p.g_col[n * 3 + j] : 0.f;
                gv[j] = g1 * wa;
                gr[j] = g2 * wa;
                ga += (g1 * (fp - np_) + g2 * (fc - nc)) * rden;
                gc += (g1 * (op - np_) + g2 * (oc - nc)) * rden;
                if (p.g_old_pts) p.g_old_pts[n * 3 + j] = g1 * wc;
                if (p.g_old_col) p.g_old_col[n * 3 + j] = g2 * wc;
            }
            if (p.g_old_cc) p.g_old_cc[n] = gc;
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            p.g_vertex_g[i * 3 + j] = gv[j];
            p.g_rgb[i * 3 + j] = gr[j];
        }
        p.g_alpha[i] = ga;
    }
}

// ---------------------------------------------------------------------------------------------
// Whole-sequence kernel: ONE cooperative launch fuses all L frames.  Every CTA is resident (cooperative
// launch), the phases of a frame are separated by a grid-wide barrier instead of a kernel boundary:
//
//   prologue   map [n,3] arrays -> float4 working map, maps(0)                        | barrier
//   frame s    P1  every map point: project, compare, 128-bit compare-and-swap minimum of (key, map index)  | barrier
//              P3  pixels, a contiguous range per CTA: merge matched map points, count the valid unmatched
//                  ones, publish the count, maps(s+1) into the other buffer set while the counts of the
//                  preceding CTAs arrive, append in row-major pixel order, last CTA writes the new size  | barrier
//   epilogue   float4 working map -> the caller's [n,3] arrays
//
// Two barriers per frame (2-3 us each) replace nine stream operations.  Inside the kernel the map and
// the per-frame maps are float4 records ({x,y,z,ccount}, {nx,ny,nz,-}, {r,g,b,-}; {vx,vy,vz,alpha}, {nx,ny,nz,valid}):
// a [n,3] fp32 array costs three 12-sector requests per warp access, a float4 record one 16-sector request,
// and these phases are bound by exactly that request rate.  The arithmetic is that of the per-frame kernels
// above (same device functions), so the map is the same bit for bit.
// ---------------------------------------------------------------------------------------------
#ifndef E2E_SEQ_NT
#define E2E_SEQ_NT 576                  // 2 CTAs/SM x 576 threads x 148 SMs >= 480*640 / 2: two pixels per thread in P3
#endif
constexpr int SEQ_NT = E2E_SEQ_NT;
constexpr int SEQ_NW = SEQ_NT / 32;
constexpr int SEQ_MAX_ITERS = 32;       // pixel sub-blocks of SEQ_NT per CTA in P3 (one flag bit each)

// 128-bit association record of a pixel: hi = (1/(c + 1e-20), dist^2) as in the two-pass kernels, lo = map index.  The
// lexicographic minimum over (hi, lo) is exactly what pass 1 + pass 2 select, so ONE compare-and-swap loop per candidate
// (atom.cas.b128, sm_90+) replaces the second pass over the candidates and its grid barrier.
struct __align__(16) Key128 {
    unsigned long long lo, hi;
};

__device__ __forceinline__ Key128 cas128(Key128 *addr, Key128 expect, Key128 val)
{
    Key128 old;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old.lo), "=l"(old.hi)
                 : "l"(expect.lo), "l"(expect.hi), "l"(val.lo), "l"(val.hi), "l"(addr)
                 : "memory");
    return old;
}

struct SeqParams {
    const float *depth, *rgb, *K, *poses;
    int L, H, W;
    float two_sigma2;
    AssocConst ac;
    float *pts, *nrm, *col, *cc;        // the caller's map arrays (read in the prologue, written in the epilogue)
    long long *n_map;
    long long capacity;
    float4 *pts4, *nrm4, *col4;         // working map
    float4 *vg4[2], *ng4[2];            // per-frame maps, two buffer sets
    Key128 *keys[2];                    // per pixel: the best candidate's (key, map index), all ones = none
    unsigned long long *counts;         // [gridDim.x] (frame + 1) << 32 | appended pixels of the CTA
    unsigned *barrier;                  // `go` word of the grid barrier (per-CTA arrival slots live at counts + 1024)
};

// The per-CTA append counts travel in ONE word together with their frame tag and nothing else is read on their strength,
// so relaxed accesses are enough (a release store would first wait for the CTA's merge stores to drain, ~1.5 us).
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#ifdef E2E_SEQ_TIMING
__device__ __forceinline__ void seq_stamp(const unsigned long long *counts, int idx)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        const_cast<unsigned long long *>(counts)[2048 + idx] = t;
    }
}
#define SEQ_STAMP(i) seq_stamp(p.counts, (i))
__device__ __forceinline__ void seq_stamp3(const unsigned long long *counts, int idx)
{
    const int which = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x / 2 ? 1 : (blockIdx.x == gridDim.x - 1 ? 2 : -1));
    if (which >= 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        const_cast<unsigned long long *>(counts)[2048 + 512 + which * 128 + idx] = t;
    }
}
#define SEQ_STAMP3(i) seq_stamp3(p.counts, (i))
#else
#define SEQ_STAMP(i)
#define SEQ_STAMP3(i)
#endif

// Grid-wide barrier (all CTAs are co-resident: cooperative launch) without a contended atomic: 296 same-address atomics
// serialise at ~27 cycles each in L2 (4 us per barrier, measured).  Every CTA release-stores the barrier's epoch into its
// own slot; warp 0 of CTA 0 polls all slots (a few per lane, loads batched), fences and release-stores the epoch into `go`; thread 0 of
// every other CTA polls `go` and fences (acquire side; the fence also drops the SM's stale L1 lines before the next phase).
// Release is cumulative over the CTA's writes because it follows bar.sync.  No sequentially-consistent fence anywhere
// (__threadfence() is MEMBAR.SC.GPU).
__device__ __forceinline__ void grid_barrier(unsigned *go, unsigned *slots, unsigned &epoch)
{
    __syncthreads();
    epoch++;
    if (blockIdx.x == 0) {
        if (threadIdx.x < 32) {
            // each lane watches every 32nd slot, eight independent loads per round trip (a serial poll of ~10 slots per
            // lane costs ~0.3 us each)
            for (unsigned b0 = 1 + threadIdx.x; b0 < gridDim.x; b0 += 32 * 8) {
                bool all;
                do {
                    all = true;
                    unsigned g[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const unsigned b = b0 + 32 * j;
                        g[j] = epoch;
                        if (b < gridDim.x) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(g[j]) : "l"(slots + b) : "memory");
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) all &= (int)(g[j] - epoch) >= 0;
                } while (!all);
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            __syncwarp();
            if (threadIdx.x == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(epoch) : "memory");   // after the fence: a release pattern
        }
    } else if (threadIdx.x == 0) {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(slots + blockIdx.x), "r"(epoch) : "memory");
        unsigned g;
        do asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(go) : "memory");
        while ((int)(g - epoch) < 0);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

// maps(f): the CTA's pixel range, two pixels per thread and trip with all six depth loads issued first.
template <int SB>
__device__ __forceinline__ void seq_maps(const SeqParams &p, int f, const Cam &c, int pix0, int chunk, int iters)
{
    constexpr int sb = SB;
    const int HW = p.H * p.W;
    const float *depth = p.depth + (size_t)f * HW;
    float4 *vg4 = p.vg4[SB], *ng4 = p.ng4[SB];
    Key128 *keys = p.keys[SB];
    for (int k0 = 0; k0 < iters; k0 += 2) {
        int i[2], y[2], x[2];
        bool on[2];
        float d[2], dr[2], dd[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int r = (k0 + e) * SEQ_NT + threadIdx.x;
            i[e] = pix0 + r;
            on[e] = (k0 + e < iters) && r < chunk && i[e] < HW;
            y[e] = on[e] ? i[e] / p.W : 0;
            x[e] = on[e] ? i[e] - y[e] * p.W : 0;
            d[e] = on[e] ? depth[i[e]] : 0.0f;
            dr[e] = (on[e] && x[e] + 1 < p.W) ? depth[i[e] + 1] : 0.0f;
            dd[e] = (on[e] && y[e] + 1 < p.H) ? depth[i[e] + p.W] : 0.0f;
        }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (!on[e]) continue;
            PixelMaps o;
            rgbd_pixel_core(d[e], dr[e], dd[e], p.H, p.W, y[e], x[e], c, p.two_sigma2, o);
            vg4[i[e]] = make_float4(o.vg[0], o.vg[1], o.vg[2], o.alpha);
            ng4[i[e]] = make_float4(o.ng[0], o.ng[1], o.ng[2], o.valid);
            keys[i[e]] = Key128{~0ull, ~0ull};
        }
    }
}

struct SeqShared {
    Cam cams[2];
    int wcnt[SEQ_MAX_ITERS * SEQ_NW];      // appended pixels per (sub-block, warp), then exclusive offsets
    long long wsum[SEQ_NW];                // look-back: per-warp sums of the preceding CTAs' counts
    long long sh_base;
};

// One frame; SB = s & 1 selects the buffer set at compile time (the pointers stay in the constant bank).
template <int SB>
__device__ __forceinline__ void seq_frame(const SeqParams &p, SeqShared &sh, int s, unsigned &target, int pix0, int chunk, int iters)
{
    constexpr int sb = SB;
    Cam *cams = sh.cams;
    int *wcnt = sh.wcnt;
    long long &sh_base = sh.sh_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int HW = p.H * p.W;
    const long long gtid = (long long)blockIdx.x * SEQ_NT + tid, gthreads = (long long)gridDim.x * SEQ_NT;
    const Cam &c = cams[sb];
    const float4 *vg4 = p.vg4[SB], *ng4 = p.ng4[SB];
    Key128 *keys = p.keys[SB];
    const long long N = *reinterpret_cast<volatile long long *>(p.n_map);
    if (tid == 32) {                                        // off the critical path: next frame's camera, next counter
        if (s + 1 < p.L) build_cam(p.K, p.poses + (size_t)(s + 1) * 16, cams[sb ^ 1]);
    }

    // ---- P1: one map point per thread and trip.  The next trip's point is loaded before this trip's gathers are used, and
    // the compare-and-swap of a candidate is only looked at one trip later, so neither round trip stalls the thread.
    {
        long long n = gtid;
        float4 nxt = (n < N) ? p.pts4[n] : make_float4(0.f, 0.f, 0.f, 0.f);
        bool pending = false;
        int ppix = 0;
        Key128 mine{0ull, 0ull}, seen{0ull, 0ull}, old{0ull, 0ull};
        // a failed attempt returns what is there; the candidate retries only while it is smaller (rare: most pixels see one
        // candidate, and the first attempt expects the initial all-ones record)
        auto resolve = [&]() {
            while (!(old.hi == seen.hi && old.lo == seen.lo)) {
                seen = old;
                if (!(mine.hi < seen.hi || (mine.hi == seen.hi && mine.lo < seen.lo))) break;
                old = cas128(keys + ppix, seen, mine);
            }
            pending = false;
        };
        for (; n < N; n += gthreads) {
            const float4 cur = nxt;
            const int pix = assoc_pixel(c, p.ac, cur.x, cur.y, cur.z);
            float4 nr, lv, ln;
            if (pix >= 0) {
                nr = p.nrm4[n];
                lv = vg4[pix];
                ln = ng4[pix];
            }
            if (n + gthreads < N) nxt = p.pts4[n + gthreads];
            if (pending) resolve();
            unsigned long long key = 0ull;
            if (pix >= 0 && assoc_key(p.ac, cur.x, cur.y, cur.z, nr.x, nr.y, nr.z, cur.w, lv.x, lv.y, lv.z, ln.x, ln.y, ln.z, key)) {
                mine = Key128{(unsigned long long)n, key};
                seen = Key128{~0ull, ~0ull};
                ppix = pix;
                old = cas128(keys + ppix, seen, mine);
                pending = true;
            }
        }
        if (pending) resolve();
    }
    grid_barrier(p.barrier, reinterpret_cast<unsigned *>(p.counts + 1024), target);
    SEQ_STAMP(2 + 3 * s);
    SEQ_STAMP(3 + 3 * s);

    // ---- P3 ------------------------------------------------------------------------------------------
    const float *rgb = p.rgb + (size_t)s * HW * 3;
    unsigned flags = 0u;                                   // bit k: this thread's pixel of sub-block k is appended
    for (int k0 = 0; k0 < iters; k0 += 2) {
        int i[2];
        bool on[2];
        unsigned n[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {                       // both index-map loads first
            const int r = (k0 + e) * SEQ_NT + tid;
            i[e] = pix0 + r;
            on[e] = (k0 + e < iters) && r < chunk && i[e] < HW;
            n[e] = on[e] ? (unsigned)keys[i[e]].lo : 0xffffffffu;
        }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            bool flag = false;
            if (n[e] != 0xffffffffu) {
                const float4 mp = p.pts4[n[e]], mn = p.nrm4[n[e]], mc = p.col4[n[e]];
                const float4 lv = vg4[i[e]], ln = ng4[i[e]];
                const float r = rgb[i[e] * 3], g = rgb[i[e] * 3 + 1], bl = rgb[i[e] * 3 + 2];
                const float cw = mp.w, a = lv.w, den = xadd(cw, a);
                p.pts4[n[e]] = make_float4(merge_val(cw, mp.x, a, lv.x, den), merge_val(cw, mp.y, a, lv.y, den),
                                           merge_val(cw, mp.z, a, lv.z, den), den);
                p.nrm4[n[e]] = make_float4(merge_val(cw, mn.x, a, ln.x, den), merge_val(cw, mn.y, a, ln.y, den),
                                           merge_val(cw, mn.z, a, ln.z, den), 0.0f);
                p.col4[n[e]] = make_float4(merge_val(cw, mc.x, a, r, den), merge_val(cw, mc.y, a, g, den),
                                           merge_val(cw, mc.z, a, bl, den), 0.0f);
            } else if (on[e]) {
                flag = ng4[i[e]].w != 0.0f;
            }
            if (k0 + e < iters) {
                const unsigned ballot = __ballot_sync(0xffffffffu, flag);
                if (flag) flags |= 1u << (k0 + e);
                if (lane == 0) wcnt[(k0 + e) * SEQ_NW + wid] = __popc(ballot);
            }
        }
    }
    __syncthreads();
    SEQ_STAMP(256 + 4 * s);
    if (wid == 0) {                                        // exclusive scan of wcnt in (sub-block, warp) order = pixel order
        int carry = 0;
        for (int b0 = 0; b0 < iters * SEQ_NW; b0 += 32) {
            const int v = (b0 + lane < iters * SEQ_NW) ? wcnt[b0 + lane] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (b0 + lane < iters * SEQ_NW) wcnt[b0 + lane] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) st_relaxed_u64(p.counts + blockIdx.x, ((unsigned long long)(s + 1) << 32) | (unsigned)carry);
    }
    __syncthreads();
    SEQ_STAMP(257 + 4 * s);
    if (s + 1 < p.L) seq_maps<SB ^ 1>(p, s + 1, cams[sb ^ 1], pix0, chunk, iters);   // independent of the map: fills the wait for the other CTAs
    SEQ_STAMP(258 + 4 * s);
    {                                                      // sum of the counts of all preceding CTAs (same frame tag): every thread
        long long before = 0;                              // looks at one of them -- one round trip instead of blockIdx.x / 32 in a row
        for (int b = tid; b < (int)blockIdx.x; b += SEQ_NT) {
            unsigned long long v;
            do v = ld_relaxed_u64(p.counts + b);
            while ((unsigned)(v >> 32) != (unsigned)(s + 1));
            before += (long long)(unsigned)v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        if (lane == 0) sh.wsum[wid] = before;
        __syncthreads();
        if (wid == 0) {
            long long t = lane < SEQ_NW ? sh.wsum[lane] : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) sh_base = N + t;
        }
    }
    __syncthreads();
    SEQ_STAMP(259 + 4 * s);
    const long long base = sh_base;
    for (int k = 0; k < iters; k++) {
        const bool flag = (flags >> k) & 1u;
        const unsigned ballot = __ballot_sync(0xffffffffu, flag);
        if (flag) {
            const int i = pix0 + k * SEQ_NT + tid;
            const long long slot = base + wcnt[k * SEQ_NW + wid] + __popc(ballot & ((1u << lane) - 1));
            if (slot < p.capacity) {
                const float4 lv = vg4[i], ln = ng4[i];
                const float r = rgb[i * 3], g = rgb[i * 3 + 1], bl = rgb[i * 3 + 2];
                p.pts4[slot] = lv;                                     // {vertex, alpha}: alpha is the new point's confidence
                p.nrm4[slot] = make_float4(ln.x, ln.y, ln.z, 0.0f);
                p.col4[slot] = make_float4(r, g, bl, 0.0f);
            }
        }
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        // total = everything before the last CTA + its own count
        const unsigned long long own = ld_relaxed_u64(p.counts + blockIdx.x);
        long long total = base + (long long)(unsigned)own;
        if (total > p.capacity) total = p.capacity;
        *reinterpret_cast<volatile long long *>(p.n_map) = total;
    }
    grid_barrier(p.barrier, reinterpret_cast<unsigned *>(p.counts + 1024), target);
    SEQ_STAMP(4 + 3 * s);
}

// The whole sequence loop of one CTA group.  Only blockIdx.x / gridDim.x are used below (pixel ranges, look-back counts, the grid
// barrier's slots): a batch of independent sequences is therefore a 2-D grid, blockIdx.y = sequence, every sequence with its own
// parameter block (own barrier word, counts and working buffers).
__device__ __forceinline__ void fusion_sequence_body(const SeqParams &p)
{
    __shared__ SeqShared sh;
    Cam *cams = sh.cams;
    const int tid = threadIdx.x;
    const int HW = p.H * p.W;
    const int chunk = (HW + gridDim.x - 1) / gridDim.x;       // contiguous pixels per CTA in P3
    const int iters = (chunk + SEQ_NT - 1) / SEQ_NT;
    const int pix0 = blockIdx.x * chunk;
    const long long gtid = (long long)blockIdx.x * SEQ_NT + tid, gthreads = (long long)gridDim.x * SEQ_NT;
    unsigned target = 0;

#ifdef E2E_SEQ_TIMING
    for (int j = 0; j < 16; j++) {
        SEQ_STAMP(1024 + j);
        grid_barrier(p.barrier, reinterpret_cast<unsigned *>(p.counts + 1024), target);
    }
    SEQ_STAMP(1024 + 16);
#endif
    SEQ_STAMP(0);
    if (tid == 0) {
        build_cam(p.K, p.poses, cams[0]);
    }
    {   // the caller's map (if any) -> working records
        const long long N0 = *reinterpret_cast<volatile long long *>(p.n_map);
        for (long long n = gtid; n < N0; n += gthreads) {
            p.pts4[n] = make_float4(p.pts[n * 3], p.pts[n * 3 + 1], p.pts[n * 3 + 2], p.cc[n]);
            p.nrm4[n] = make_float4(p.nrm[n * 3], p.nrm[n * 3 + 1], p.nrm[n * 3 + 2], 0.0f);
            p.col4[n] = make_float4(p.col[n * 3], p.col[n * 3 + 1], p.col[n * 3 + 2], 0.0f);
        }
    }
    __syncthreads();
    seq_maps<0>(p, 0, cams[0], pix0, chunk, iters);
    grid_barrier(p.barrier, reinterpret_cast<unsigned *>(p.counts + 1024), target);
    SEQ_STAMP(1);

    for (int s = 0; s < p.L; s++) {
        if (s & 1) seq_frame<1>(p, sh, s, target, pix0, chunk, iters);
        else seq_frame<0>(p, sh, s, target, pix0, chunk, iters);
    }

    // ---- epilogue: working records -> the caller's [n,3] arrays -----------------------------------------
    const long long N = *reinterpret_cast<volatile long long *>(p.n_map);
    for (long long n = gtid; n < N; n += gthreads) {
        const float4 a = p.pts4[n], b = p.nrm4[n], cl = p.col4[n];
        p.pts[n * 3] = a.x; p.pts[n * 3 + 1] = a.y; p.pts[n * 3 + 2] = a.z;
        p.nrm[n * 3] = b.x; p.nrm[n * 3 + 1] = b.y; p.nrm[n * 3 + 2] = b.z;
        p.col[n * 3] = cl.x; p.col[n * 3 + 1] = cl.y; p.col[n * 3 + 2] = cl.z;
        p.cc[n] = a.w;
    }
}

__global__ void __launch_bounds__(SEQ_NT, 2) fusion_sequence_kernel(const SeqParams p) { fusion_sequence_body(p); }

// B independent sequences in ONE cooperative launch: grid (CTAs per sequence, B).  A single sequence is bound by dependent round
// trips and two grid barriers per frame (8 % of DRAM peak, 30 % issue slots); sequences that share the SMs fill each other's
// bubbles.  Parameter blocks live in global memory (one per sequence).
__global__ void __launch_bounds__(SEQ_NT, 2) fusion_sequence_batch_kernel(const SeqParams *pp)
{
    __shared__ SeqParams sp;
    static_assert(sizeof(SeqParams) % 4 == 0 && sizeof(SeqParams) / 4 <= SEQ_NT, "parameter block is staged by one pass of the CTA");
    if (threadIdx.x < sizeof(SeqParams) / 4) reinterpret_cast<unsigned *>(&sp)[threadIdx.x] = reinterpret_cast<const unsigned *>(pp + blockIdx.y)[threadIdx.x];
    __syncthreads();
    const SeqParams p = sp;                 // a private copy: the fields the loop uses live in registers, like kernel parameters
    fusion_sequence_body(p);
}

static inline int grid_for(long long n)
{
    long long b = (n + FU_NT - 1) / FU_NT;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace e2e

using namespace e2e;

extern "C" {

int e2e_rgbd_maps(const float *depth, const float *rgb, const float *K, const float *pose, int H, int W, float sigma,
                  float *vertex_g, float *normal_g, float *alpha, unsigned char *valid, void *stream)
{
    (void)rgb;   // colours are consumed by the merge / append step as they are
    E2E_REQUIRE(depth && K && pose && vertex_g && normal_g && alpha && valid && H > 0 && W > 0, "rgbd_maps: bad arguments");
    E2E_REQUIRE(sigma != 0.0f, "sigma must be non-zero");
    rgbd_maps_kernel<<<grid_for((long long)H * W), FU_NT, 0, (cudaStream_t)stream>>>(depth, K, pose, H, W, 2.0f * sigma * sigma,
                                                                                 vertex_g, normal_g, alpha, valid);
    count_launch();
    return finish_launch("rgbd_maps");
}

int e2e_rgbd_maps_bwd(const float *depth, const float *K, const float *pose, int H, int W, float sigma,
                      const float *grad_vertex_g, const float *grad_normal_g, const float *grad_alpha,
                      float *grad_depth, void *stream)
{
    E2E_REQUIRE(depth && K && pose && grad_depth && H > 0 && W > 0, "rgbd_maps_bwd: bad arguments");
    if (grad_normal_g) {
        set_error("rgbd_maps_bwd: normals are non-differentiable outputs in this implementation");
        return E2E_ERR_UNSUPPORTED;
    }
    rgbd_maps_bwd_kernel<<<grid_for((long long)H * W), FU_NT, 0, (cudaStream_t)stream>>>(depth, K, pose, H, W, 2.0f * sigma * sigma,
                                                                                     grad_vertex_g, grad_alpha, grad_depth);
    count_launch();
    return finish_launch("rgbd_maps_bwd");
}

int e2e_fusion_associate(const float *map_points, const float *map_normals, const float *map_ccount,
                         const long long *n_map, long long n_upper,
                         const float *K, const float *pose, const float *vertex_g, const float *normal_g,
                         int H, int W, float dist_th, float dot_th,
                         unsigned long long *keys, int *candidates, long long *index_map, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    E2E_REQUIRE(n_map && K && pose && vertex_g && normal_g && keys && index_map && H > 0 && W > 0 && n_upper >= 0,
                "fusion_associate: bad arguments");
    E2E_REQUIRE(n_upper == 0 || candidates, "fusion_associate: candidates scratch (int32[n_upper]) is null");
    E2E_REQUIRE((long long)H * W < (1ll << 31), "fusion_associate: image too large");
    E2E_REQUIRE(n_upper == 0 || (map_points && map_normals && map_ccount), "fusion_associate: null map");
    // keys = all ones (maximum); index_map = all ones = -1 as int64 = maximum as uint64
    cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * (size_t)H * W, st);
    cudaMemsetAsync(index_map, 0xff, sizeof(long long) * (size_t)H * W, st);
    if (n_upper == 0) return finish_launch("fusion_associate");
    AssocParams p;
    p.pts = map_points; p.nrm = map_normals; p.cc = map_ccount; p.n_map = n_map; p.n_upper = n_upper;
    p.K = K; p.pose = pose; p.vertex_g = vertex_g; p.normal_g = normal_g; p.H = H; p.W = W;
    p.dist_th = dist_th; p.dot_th = dot_th;
    p.u_hi = (float)((double)W - 0.999); p.v_hi = (float)((double)H - 0.999);
    p.keys = keys; p.index_map = (unsigned long long *)index_map; p.cand = candidates;
    const int grid = grid_for(n_upper);
    associate_pass1_kernel<<<grid, FU_NT, 0, st>>>(p);
    associate_pass2_kernel<<<grid, FU_NT, 0, st>>>(p);
    count_launch(2);
    return finish_launch("fusion_associate");
}

size_t e2e_fusion_active_points_workspace_bytes(long long n)
{
    const size_t nchunks = (size_t)((n + FU_NT - 1) / FU_NT);
    return (size_t)(n > 0 ? n : 0) * sizeof(int) + nchunks * sizeof(int) + 256;
}

int e2e_fusion_active_points(const float *map_points, long long n, const float *K, const float *pose, int H, int W,
                             long long batch_index, long long *rows, long long *n_active,
                             void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    E2E_REQUIRE(K && pose && n_active && H > 0 && W > 0 && n >= 0, "fusion_active_points: bad arguments");
    E2E_REQUIRE((long long)H * W < (1ll << 31), "fusion_active_points: image too large");
    if (n == 0) {
        cudaMemsetAsync(n_active, 0, sizeof(long long), st);
        return finish_launch("fusion_active_points");
    }
    E2E_REQUIRE(map_points && rows, "fusion_active_points: null map / rows");
    const long long nchunks = (n + FU_NT - 1) / FU_NT;
    E2E_REQUIRE(nchunks < (1ll << 31), "fusion_active_points: map too large");
    E2E_REQUIRE(workspace && workspace_bytes >= (size_t)n * sizeof(int) + (size_t)nchunks * sizeof(int), "fusion_active_points: workspace too small");
    ActiveParams p;
    p.pts = map_points; p.n = n; p.K = K; p.pose = pose; p.H = H; p.W = W;
    p.u_hi = (float)((double)W - 0.999); p.v_hi = (float)((double)H - 0.999);
    p.batch = batch_index;
    p.pix = (int *)workspace; p.chunk_counts = p.pix + n; p.rows = rows;
    active_count_kernel<<<(unsigned)nchunks, FU_NT, 0, st>>>(p);
    fuse_scan_kernel<<<1, 1024, 0, st>>>(p.chunk_counts, (int)nchunks, nullptr, n_active);
    active_write_kernel<<<(unsigned)nchunks, FU_NT, 0, st>>>(p);
    count_launch(3);
    return finish_launch("fusion_active_points");
}

size_t e2e_fusion_workspace_bytes(int H, int W)
{
    const size_t nchunks = ((size_t)H * W + SCAN_CHUNK - 1) / SCAN_CHUNK;
    return nchunks * sizeof(int) + 256;
}

int e2e_fusion_merge_append(float *map_points, float *map_normals, float *map_colors, float *map_ccount,
                            const long long *n_map, long long capacity,
                            const float *vertex_g, const float *normal_g, const float *rgb, const float *alpha,
                            const unsigned char *valid, const long long *index_map, int H, int W,
                            long long *append_slot, long long *n_out,
                            void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    E2E_REQUIRE(map_points && map_normals && map_colors && map_ccount && n_map && vertex_g && normal_g && rgb && alpha && valid &&
                index_map && n_out && H > 0 && W > 0, "fusion_merge_append: bad arguments");
    const int nchunks = (int)(((long long)H * W + SCAN_CHUNK - 1) / SCAN_CHUNK);
    E2E_REQUIRE(workspace && workspace_bytes >= nchunks * sizeof(int), "fusion_merge_append: workspace too small");
    FuseParams p;
    p.pts = map_points; p.nrm = map_normals; p.col = map_colors; p.cc = map_ccount; p.n_map = n_map; p.capacity = capacity;
    p.vertex_g = vertex_g; p.normal_g = normal_g; p.rgb = rgb; p.alpha = alpha; p.valid = valid; p.index_map = index_map;
    p.H = H; p.W = W; p.append_slot = append_slot; p.n_out = n_out; p.chunk_counts = (int *)workspace; p.nchunks = nchunks;
    fuse_merge_count_kernel<<<nchunks, FU_NT, 0, st>>>(p);
    fuse_scan_kernel<<<1, 1024, 0, st>>>(p.chunk_counts, nchunks, n_map, n_out);
    fuse_append_kernel<<<nchunks, FU_NT, 0, st>>>(p);
    count_launch(3);
    return finish_launch("fusion_merge_append");
}

int e2e_fusion_merge_append_bwd(const float *grad_points, const float *grad_colors, const float *grad_ccount,
                                const float *old_points, const float *old_colors, const float *old_ccount,
                                const float *vertex_g, const float *rgb, const float *alpha,
                                const long long *index_map, const long long *append_slot, int H, int W,
                                float *grad_vertex_g, float *grad_rgb, float *grad_alpha,
                                float *grad_old_points, float *grad_old_colors, float *grad_old_ccount, void *stream)
{
    E2E_REQUIRE(vertex_g && rgb && alpha && index_map && append_slot && grad_vertex_g && grad_rgb && grad_alpha && H > 0 && W > 0,
                "fusion_merge_append_bwd: bad arguments");
    FuseBwdParams p;
    p.g_pts = grad_points; p.g_col = grad_colors; p.g_cc = grad_ccount;
    p.old_pts = old_points; p.old_col = old_colors; p.old_cc = old_ccount;
    p.vertex_g = vertex_g; p.rgb = rgb; p.alpha = alpha; p.index_map = index_map; p.append_slot = append_slot;
    p.H = H; p.W = W; p.g_vertex_g = grad_vertex_g; p.g_rgb = grad_rgb; p.g_alpha = grad_alpha;
    p.g_old_pts = grad_old_points; p.g_old_col = grad_old_colors; p.g_old_cc = grad_old_ccount;
    fuse_bwd_kernel<<<grid_for((long long)H * W), FU_NT, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return finish_launch("fusion_merge_append_bwd");
}

/* ---------------------------------------------------------------------------------------------
 * Whole-sequence fusion with known poses (slam/custom_slam.py:26-34 / PointFusion.forward with odom="gt", no
 * autograd): the frame loop runs here, so the host pays one call per SEQUENCE instead of ~15 launches, memsets and
 * allocations per frame from Python (which bound the per-frame time at ~130 us on the host, more than the kernels).
 * --------------------------------------------------------------------------------------------- */
static size_t align256(size_t n) { return (n + 255) / 256 * 256; }

// Grid of the whole-sequence kernel: every CTA must be resident (cooperative launch).  0 = not available.
static int sequence_grid()
{
    static int grid_dev[64];                       // per device, 0 = not yet known
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && grid_dev[dev & 63] > 0) return grid_dev[dev & 63];
    int &grid = grid_dev[dev & 63];
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fusion_sequence_kernel, SEQ_NT, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;                       // not cached: a later call on a working context may succeed
    }
    grid = coop ? sms * per_sm : 0;
    return grid;
}

static size_t sequence_loop_bytes(size_t hw, long long capacity, int H, int W)
{
    return align256(hw * 12) * 2 + align256(hw * 4) + align256(hw) + align256(hw * 8) * 2 + align256((size_t)capacity * 4) +
           align256(e2e_fusion_workspace_bytes(H, W)) + 256;
}

static size_t sequence_coop_bytes(size_t hw, long long capacity)
{
    // working map (3 float4 records per point), two buffer sets of {vertex+alpha, normal+valid, association record},
    // per-CTA counts / barrier slots / timing stamps, barrier word
    return 3 * align256((size_t)capacity * 16) + 2 * align256(hw * 16) * 3 + align256(8 * 4096) + 256;
}

size_t e2e_fusion_sequence_workspace_bytes(int H, int W, long long capacity)
{
    const size_t hw = (size_t)H * W;
    const size_t a = sequence_loop_bytes(hw, capacity, H, W), b = sequence_coop_bytes(hw, capacity);
    return a > b ? a : b;
}

// The per-frame loop over the stand-alone entry points (also the reference the cooperative kernel is tested against).
static int fusion_sequence_loop(const float *depth, const float *rgb, const float *K, const float *poses, int L, int H, int W,
                                float sigma, float dist_th, float dot_th,
                                float *map_points, float *map_normals, float *map_colors, float *map_ccount,
                                long long *n_map, long long n_upper, long long capacity, void *workspace, void *stream)
{
    const size_t hw = (size_t)H * W;
    unsigned char *w = (unsigned char *)workspace;
    float *vg = (float *)w;                                 w += align256(hw * 12);
    float *ng = (float *)w;                                 w += align256(hw * 12);
    float *alpha = (float *)w;                              w += align256(hw * 4);
    unsigned char *valid = w;                               w += align256(hw);
    unsigned long long *keys = (unsigned long long *)w;     w += align256(hw * 8);
    long long *index_map = (long long *)w;                  w += align256(hw * 8);
    int *cand = (int *)w;                                   w += align256((size_t)capacity * 4);
    void *ws = w;
    const size_t ws_bytes = e2e_fusion_workspace_bytes(H, W);
    for (int s = 0; s < L; s++) {
        const float *pose = poses + (size_t)s * 16;
        if (int rc = e2e_rgbd_maps(depth + (size_t)s * hw, nullptr, K, pose, H, W, sigma, vg, ng, alpha, valid, stream)) return rc;
        const long long upper = n_upper + (long long)s * H * W;
        if (int rc = e2e_fusion_associate(map_points, map_normals, map_ccount, n_map, upper, K, pose, vg, ng, H, W, dist_th, dot_th,
                                          keys, cand, index_map, stream)) return rc;
        // the new point count goes to the scratch slot n_map[1] (fuse_append still needs the old count), then replaces n_map[0]
        if (int rc = e2e_fusion_merge_append(map_points, map_normals, map_colors, map_ccount, n_map, capacity, vg, ng,
                                             rgb + (size_t)s * hw * 3, alpha, valid, index_map, H, W, nullptr, n_map + 1,
                                             ws, ws_bytes, stream)) return rc;
        if (cudaMemcpyAsync(n_map, n_map + 1, sizeof(long long), cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess)
            return finish_launch("fusion_sequence: n_map update");
    }
    return 0;
}

static SeqParams make_seq_params(const float *depth, const float *rgb, const float *K, const float *poses, int L, int H, int W,
                                 float sigma, float dist_th, float dot_th, float *map_points, float *map_normals, float *map_colors,
                                 float *map_ccount, long long *n_map, long long capacity, void *workspace)
{
    const size_t hw = (size_t)H * W;
    SeqParams p;
    p.depth = depth; p.rgb = rgb; p.K = K; p.poses = poses; p.L = L; p.H = H; p.W = W;
    p.two_sigma2 = 2.0f * sigma * sigma;
    p.ac = AssocConst{H, W, dist_th, dot_th, (float)((double)W - 0.999), (float)((double)H - 0.999)};
    p.pts = map_points; p.nrm = map_normals; p.col = map_colors; p.cc = map_ccount; p.n_map = n_map; p.capacity = capacity;
    unsigned char *w = (unsigned char *)workspace;
    p.pts4 = (float4 *)w;                           w += align256((size_t)capacity * 16);
    p.nrm4 = (float4 *)w;                           w += align256((size_t)capacity * 16);
    p.col4 = (float4 *)w;                           w += align256((size_t)capacity * 16);
    for (int b = 0; b < 2; b++) {
        p.vg4[b] = (float4 *)w;                     w += align256(hw * 16);
        p.ng4[b] = (float4 *)w;                     w += align256(hw * 16);
        p.keys[b] = (Key128 *)w;                    w += align256(hw * 16);
    }
    p.counts = (unsigned long long *)w;             w += align256(8 * 4096);
    p.barrier = (unsigned *)w;
    return p;
}

int e2e_fusion_sequence(const float *depth, const float *rgb, const float *K, const float *poses, int L, int H, int W,
                        float sigma, float dist_th, float dot_th,
                        float *map_points, float *map_normals, float *map_colors, float *map_ccount,
                        long long *n_map, long long n_upper, long long capacity,
                        void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(depth && rgb && K && poses && map_points && map_normals && map_colors && map_ccount && n_map && workspace,
                "fusion_sequence: null argument");
    E2E_REQUIRE(L >= 0 && H > 0 && W > 0 && n_upper >= 0, "fusion_sequence: bad sizes");
    E2E_REQUIRE(capacity >= n_upper + (long long)L * H * W, "fusion_sequence: capacity must be >= n_upper + L*H*W");
    E2E_REQUIRE(workspace_bytes >= e2e_fusion_sequence_workspace_bytes(H, W, capacity), "fusion_sequence: workspace too small");
    E2E_REQUIRE(sigma != 0.0f, "sigma must be non-zero");
    E2E_REQUIRE((long long)H * W < (1ll << 31), "fusion_sequence: image too large");
    if (L == 0) return 0;
    const size_t hw = (size_t)H * W;
    // E2E_FUSION_SEQUENCE=loop forces the per-frame launches (A/B timing, debugging)
    static const bool force_loop = [] { const char *e = getenv("E2E_FUSION_SEQUENCE"); return e && e[0] == 'l'; }();
    const int grid = force_loop ? 0 : sequence_grid();
    const long long chunk = grid ? ((long long)hw + grid - 1) / grid : 0;
    const bool coop = grid > 0 && grid <= 2048 && (chunk + SEQ_NT - 1) / SEQ_NT <= SEQ_MAX_ITERS && capacity < 0xffffffffll;
    if (!coop)
        return fusion_sequence_loop(depth, rgb, K, poses, L, H, W, sigma, dist_th, dot_th, map_points, map_normals, map_colors,
                                    map_ccount, n_map, n_upper, capacity, workspace, stream);
    SeqParams p = make_seq_params(depth, rgb, K, poses, L, H, W, sigma, dist_th, dot_th, map_points, map_normals, map_colors, map_ccount,
                                  n_map, capacity, workspace);
    cudaStream_t st = (cudaStream_t)stream;
    // frame tags of the counts start at 1, the barrier counts arrivals from 0
    if (cudaMemsetAsync(p.counts, 0, align256(8 * 4096) + 256, st) != cudaSuccess) return finish_launch("fusion_sequence: memset");
    void *args[] = {(void *)&p};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)fusion_sequence_kernel, dim3(grid), dim3(SEQ_NT), args, 0, st);
    if (e == cudaErrorCooperativeLaunchTooLarge) {      // the GPU is shared (other resident kernels): per-frame launches instead
        cudaGetLastError();
        return fusion_sequence_loop(depth, rgb, K, poses, L, H, W, sigma, dist_th, dot_th, map_points, map_normals, map_colors,
                                    map_ccount, n_map, n_upper, capacity, workspace, stream);
    }
    if (e != cudaSuccess) {
        set_error("fusion_sequence: cooperative launch failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    count_launch();
    return finish_launch("fusion_sequence");
}

size_t e2e_fusion_sequence_batch_workspace_bytes(int B, int H, int W, long long capacity)
{
    return (size_t)(B > 0 ? B : 0) * (align256(e2e_fusion_sequence_workspace_bytes(H, W, capacity)) + align256(sizeof(SeqParams))) + 256;
}

int e2e_fusion_sequence_batch(const float *depth, const float *rgb, const float *K, const float *poses, int B, int L, int H, int W,
                              float sigma, float dist_th, float dot_th,
                              float *map_points, float *map_normals, float *map_colors, float *map_ccount,
                              long long *n_map, long long capacity, void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(depth && rgb && K && poses && map_points && map_normals && map_colors && map_ccount && n_map && workspace,
                "fusion_sequence_batch: null argument");
    E2E_REQUIRE(B >= 1 && L >= 0 && H > 0 && W > 0, "fusion_sequence_batch: bad sizes");
    E2E_REQUIRE(capacity >= (long long)L * H * W, "fusion_sequence_batch: capacity must be >= L*H*W");
    E2E_REQUIRE(workspace_bytes >= e2e_fusion_sequence_batch_workspace_bytes(B, H, W, capacity), "fusion_sequence_batch: workspace too small");
    E2E_REQUIRE(sigma != 0.0f, "sigma must be non-zero");
    E2E_REQUIRE((long long)H * W < (1ll << 31), "fusion_sequence_batch: image too large");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)H * W;
    const size_t per_ws = align256(e2e_fusion_sequence_workspace_bytes(H, W, capacity));
    const int total = sequence_grid();
    // as many sequences per launch as still leave every sequence enough CTAs for the kernel's pixel sub-block limit
    int per_launch = B;
    while (per_launch > 1) {
        const int g = total / per_launch;
        if (g >= 1 && ((long long)((hw + g - 1) / g) + SEQ_NT - 1) / SEQ_NT <= SEQ_MAX_ITERS) break;
        per_launch--;
    }
    const int g1 = total > 0 ? total / per_launch : 0;
    const bool coop = total > 0 && g1 >= 1 && g1 <= 2048 && capacity < 0xffffffffll &&
                      ((long long)((hw + g1 - 1) / g1) + SEQ_NT - 1) / SEQ_NT <= SEQ_MAX_ITERS;
    unsigned char *w = (unsigned char *)workspace;
    SeqParams *dev_params = (SeqParams *)(w + (size_t)B * per_ws);
    const size_t pstride = align256(sizeof(SeqParams));
    for (int b0 = 0; b0 < B; b0 += per_launch) {
        const int nb = (B - b0 < per_launch) ? B - b0 : per_launch;
        for (int b = b0; b < b0 + nb; b++) {
            const float *d_b = depth + (size_t)b * L * hw, *rgb_b = rgb + (size_t)b * L * hw * 3, *K_b = K + (size_t)b * 16, *po_b = poses + (size_t)b * L * 16;
            float *pt_b = map_points + (size_t)b * capacity * 3, *nr_b = map_normals + (size_t)b * capacity * 3,
                  *co_b = map_colors + (size_t)b * capacity * 3, *cc_b = map_ccount + (size_t)b * capacity;
            void *ws_b = w + (size_t)b * per_ws;
            if (!coop) {            // no cooperative launch available: sequences one after another through the single-sequence entry
                if (int rc = e2e_fusion_sequence(d_b, rgb_b, K_b, po_b, L, H, W, sigma, dist_th, dot_th, pt_b, nr_b, co_b, cc_b, n_map + 2 * b, 0,
                                                 capacity, ws_b, per_ws, stream)) return rc;
                continue;
            }
            const SeqParams p = make_seq_params(d_b, rgb_b, K_b, po_b, L, H, W, sigma, dist_th, dot_th, pt_b, nr_b, co_b, cc_b, n_map + 2 * b,
                                                capacity, ws_b);
            if (cudaMemsetAsync(p.counts, 0, align256(8 * 4096) + 256, st) != cudaSuccess) return finish_launch("fusion_sequence_batch: memset");
            if (cudaMemcpyAsync((unsigned char *)dev_params + (size_t)(b - b0) * sizeof(SeqParams), &p, sizeof(SeqParams), cudaMemcpyHostToDevice, st) !=
                cudaSuccess) return finish_launch("fusion_sequence_batch: parameter upload");
        }
        if (!coop) continue;
        (void)pstride;
        const SeqParams *pp = dev_params;
        void *args[] = {(void *)&pp};
        const cudaError_t e = cudaLaunchCooperativeKernel((const void *)fusion_sequence_batch_kernel, dim3(total / nb, nb), dim3(SEQ_NT), args, 0, st);
        if (e != cudaSuccess) {
            set_error("fusion_sequence_batch: cooperative launch failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return (int)e;
        }
        count_launch();
        // the parameter blocks of this wave are read by the running kernel: the next wave's upload must wait for it
        if (b0 + nb < B && cudaStreamSynchronize(st) != cudaSuccess) return finish_launch("fusion_sequence_batch: sync between waves");
    }
    return finish_launch("fusion_sequence_batch");
}

}  // extern "C"

