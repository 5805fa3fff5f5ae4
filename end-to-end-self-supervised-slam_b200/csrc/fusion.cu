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

// masked local vertex of pixel (y, x):  ((u*ifx + icx) * d, (v*ify + icy) * d, d) * [d > 0]
__device__ __forceinline__ void local_vertex(const float *depth, int W, int y, int x, const Cam &c, float V[3], float &m)
{
    const float d = depth[y * W + x];
    m = (d > 0.0f) ? 1.0f : 0.0f;
    V[0] = xmul(xmul(xadd(xmul((float)x, c.ifx), c.icx), d), m);
    V[1] = xmul(xmul(xadd(xmul((float)y, c.ify), c.icy), d), m);
    V[2] = xmul(d, m);
}

// ---------------------------------------------------------------------------------------------
// e2e_rgbd_maps
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FU_NT) rgbd_maps_kernel(const float *depth, const float *K, const float *pose, int H, int W,
                                                          float two_sigma2, float *vertex_g, float *normal_g, float *alpha,
                                                          unsigned char *valid)
{
    __shared__ Cam c;
    stage_cam(K, pose, &c);
    const int HW = H * W;
    for (int i = blockIdx.x * FU_NT + threadIdx.x; i < HW; i += gridDim.x * FU_NT) {
        const int y = i / W, x = i - y * W;
        float V[3], Vr[3], Vd[3], m, mr, md;
        local_vertex(depth, W, y, x, c, V, m);
        float dh[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f};
        if (x + 1 < W) {
            local_vertex(depth, W, y, x + 1, c, Vr, mr);
#pragma unroll
            for (int k = 0; k < 3; k++) dh[k] = xsub(Vr[k], V[k]);
        }
        if (y + 1 < H) {
            local_vertex(depth, W, y + 1, x, c, Vd, md);
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
            vertex_g[i * 3 + k] = xmul(vg, m);
            normal_g[i * 3 + k] = dot3_lr(c.R[k * 3], c.R[k * 3 + 1], c.R[k * 3 + 2], N[0], N[1], N[2]);
        }
        // alpha = exp(-(X^2 + Y^2) / (2 sigma^2)): the float32 argument is exponentiated in double and
        // rounded once, so host (numpy) and device agree on the bits of the confidence counts.
        const float arg = -xdiv(xadd(xmul(V[0], V[0]), xmul(V[1], V[1])), two_sigma2);
        alpha[i] = (float)exp((double)arg);
        valid[i] = (m != 0.0f) ? 1 : 0;
    }
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

// Pass 1: every map point is projected into the live frame; candidates (in the frustum, close in space, similar
// normal) race for their pixel with a 64-bit atomicMin of the key (1/(c+1e-20), dist^2) and remember their pixel in
// cand[n] (-1 = not a candidate).  All per-point loads (point, normal, confidence) are issued up front, so the
// dependent chain is two memory round trips (point data -> live-frame gather) instead of four.
__global__ void __launch_bounds__(FU_NT) associate_pass1_kernel(const AssocParams p)
{
    __shared__ Cam c;
    stage_cam(p.K, p.pose, &c);
    const long long N = *p.n_map;
    for (long long n = (long long)blockIdx.x * FU_NT + threadIdx.x; n < N; n += (long long)gridDim.x * FU_NT) {
        const float px = p.pts[n * 3], py = p.pts[n * 3 + 1], pz = p.pts[n * 3 + 2];
        const float nx = p.nrm[n * 3], ny = p.nrm[n * 3 + 1], nz = p.nrm[n * 3 + 2];
        const float cc = p.cc[n];
        int cand = -1;
        // 1. active map points: into the live camera, in front, inside the frustum, round to a pixel
        const float qx = xadd(dot3_lr(c.Ri[0], c.Ri[1], c.Ri[2], px, py, pz), c.ti[0]);
        const float qy = xadd(dot3_lr(c.Ri[3], c.Ri[4], c.Ri[5], px, py, pz), c.ti[1]);
        const float qz = xadd(dot3_lr(c.Ri[6], c.Ri[7], c.Ri[8], px, py, pz), c.ti[2]);
        if (qz > 0.0f) {
            const float h0 = xadd(dot3_lr(c.K[0], c.K[1], c.K[2], qx, qy, qz), c.K[3]);
            const float h1 = xadd(dot3_lr(c.K[4], c.K[5], c.K[6], qx, qy, qz), c.K[7]);
            const float h2 = xadd(dot3_lr(c.K[8], c.K[9], c.K[10], qx, qy, qz), c.K[11]);
            const float u = xdiv(h0, h2), v = xdiv(h1, h2);
            if (u > -1e-3f && u < p.u_hi && v > -1e-3f && v < p.v_hi) {
                int w = __float2int_rn(u), h = __float2int_rn(v);       // round half to even, like torch.round
                w = min(max(w, 0), p.W - 1);
                h = min(max(h, 0), p.H - 1);
                const int pix = h * p.W + w;
                // 2. similar: close in space, similar normal (both live-frame gathers issued together)
                const float vx = p.vertex_g[pix * 3], vy = p.vertex_g[pix * 3 + 1], vz = p.vertex_g[pix * 3 + 2];
                const float gx = p.normal_g[pix * 3], gy = p.normal_g[pix * 3 + 1], gz = p.normal_g[pix * 3 + 2];
                const float dx = xsub(vx, px), dy = xsub(vy, py), dz = xsub(vz, pz);
                const float dist2 = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
                const float dot = dot3_lr(gx, gy, gz, nx, ny, nz);
                if (__fsqrt_rn(dist2) < p.dist_th && dot > p.dot_th) {
                    // 3. best unique: lexicographic minimum of (1/(c + 1e-20), dist^2, n) per pixel
                    const float inv_c = xdiv(1.0f, xadd(cc, 1e-20f));
                    const unsigned long long key = ((unsigned long long)__float_as_uint(inv_c) << 32) | (unsigned long long)__float_as_uint(dist2);
                    atomicMin(p.keys + pix, key);
                    cand = pix;
                }
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
        const float dx = xsub(p.vertex_g[pix * 3], px), dy = xsub(p.vertex_g[pix * 3 + 1], py), dz = xsub(p.vertex_g[pix * 3 + 2], pz);
        const float dist2 = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
        const float inv_c = xdiv(1.0f, xadd(p.cc[n], 1e-20f));
        const unsigned long long key = ((unsigned long long)__float_as_uint(inv_c) << 32) | (unsigned long long)__float_as_uint(dist2);
        if (p.keys[pix] == key) atomicMin(p.index_map + pix, (unsigned long long)n);
    }
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
            const float c = p.cc[n], a = p.alpha[i];
            const float den = xadd(c, a);
#pragma unroll
            for (int j = 0; j < 3; j++) {       // (c*old + a*new) / (c + a), one rounding per operation
                p.pts[n * 3 + j] = xdiv(xadd(xmul(c, p.pts[n * 3 + j]), xmul(a, p.vertex_g[i * 3 + j])), den);
                p.nrm[n * 3 + j] = xdiv(xadd(xmul(c, p.nrm[n * 3 + j]), xmul(a, p.normal_g[i * 3 + j])), den);
                p.col[n * 3 + j] = xdiv(xadd(xmul(c, p.col[n * 3 + j]), xmul(a, p.rgb[i * 3 + j])), den);
            }
            p.cc[n] = den;
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
    if (threadIdx.x == 0) n_out[0] = n_map[0] + (long long)carry;
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

size_t e2e_fusion_sequence_workspace_bytes(int H, int W, long long capacity)
{
    const size_t hw = (size_t)H * W;
    return align256(hw * 12) * 2 + align256(hw * 4) + align256(hw) + align256(hw * 8) * 2 + align256((size_t)capacity * 4) +
           align256(e2e_fusion_workspace_bytes(H, W)) + 256;
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

}  // extern "C"

