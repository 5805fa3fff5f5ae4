// Fused inverse-warp + SSIM/L1 photometric loss, forward and backward, plus the stand-alone
// SSIM / photometric-loss kernels that share the same tile code (sm_100a).
//
// One CTA owns a TH x TW tile of target pixels.
//   forward : phase 1 fills shared-memory planes x = syn*valid, y = tgt*valid for the tile plus a
//             1-pixel halo (back-project -> project -> bilinear gather, all in registers), phase 2
//             evaluates the 3x3 SSIM statistics from shared memory in vertical strips and writes the
//             loss map / block partial sums.  No intermediate ever reaches HBM.
//   backward: phase A refills x, y for a 2-pixel halo, phase B recomputes the SSIM statistics at every
//             centre of the 1-pixel halo and stores three coefficient planes per channel, phase C
//             gathers them (adjoint of reflect-pad + box filter), applies the sampler / projection
//             chain rule and emits grad_depth (one owner thread per pixel, no atomics), grad_src
//             (red.global.add) and per-CTA grad_P partials (deterministic second pass).
//
// Forward arithmetic follows oracle/warp_photo_oracle.c operation for operation, which reproduces the
// reference (view_synthesis.py:34-78, F.grid_sample, losses.py:23-37, 111-115) bit for bit.
//
// Division strategy.  The reference divides by 9 (avg_pool2d), 3 (channel mean) and W-1 / H-1 with IEEE
// division.  x/d is evaluated as q = x*c, r = fma(-d,q,x), q' = fma(r,c,q) (c = RN(1/d)), which equals
// RN(x/d) for every x with |x| in [2^-122, 2^127) or x == 0 once d has passed the exhaustive mantissa
// check (e2e_prepare_divisor).  Instead of guarding each of the ~20 divisions per pixel, phase 1 checks
// the VALUES it stores (each must be 0 or have 2^-40 <= |v| <= 2^40, which bounds every window sum and
// product away from the failing range) and the CTA votes once; a CTA that sees anything else (denormal
// garbage, inf, NaN) takes the IEEE-division code path.  Either way the result is the reference's bits.
#include <cstdlib>

#include "warp_photo_common.cuh"

namespace e2e {

// ================================================================================================
// Forward kernel
// ================================================================================================
template <bool IEEE, int CK, int TH, int TW, int NT>
__device__ __forceinline__ void fwd_phase2(const WPParams &p, const float *sx, const float *sy, float *red,
                                           int b, int ch0, int ty0, int tx0)
{
    constexpr int RH = TH + 2, RP = TW + 2, NP = TH * TW / NT;
    const int H = p.H, W = p.W;
    const int col = threadIdx.x % TW, rg = threadIdx.x / TW;
    const int x = tx0 + col;
    float ssum[NP], lsum[NP];
#pragma unroll
    for (int ch = 0; ch < CK; ch++) {
        float S[NP][5];
        const float *wx = sx + (ch * RH + rg * NP) * RP + col;
        const float *wy = sy + (ch * RH + rg * NP) * RP + col;
        window_sums<NP, RP>(wx, wy, S);
#pragma unroll
        for (int pp = 0; pp < NP; pp++) {
            SsimVals v;
            ssim_finish<IEEE>(S[pp], v);
            const float cx = wx[(pp + 1) * RP + 1], cy = wy[(pp + 1) * RP + 1];
            const float l1 = fabsf(xsub(cy, cx));                                // losses.py:112
            const int y = ty0 + rg * NP + pp;
            if (p.ssim && y < H && x < W)
                p.ssim[(((long long)b * p.C + ch0 + ch) * H + y) * W + x] = v.s;
            if (ch == 0) { ssum[pp] = v.s; lsum[pp] = l1; }
            else { ssum[pp] = xadd(ssum[pp], v.s); lsum[pp] = xadd(lsum[pp], l1); }
        }
    }
    if (CK == 3 && (p.loss_map || p.partial)) {
        float acc = 0.f;
#pragma unroll
        for (int pp = 0; pp < NP; pp++) {
            const int y = ty0 + rg * NP + pp;
            if (y < H && x < W) {
                const float r3 = 1.0f / 3.0f;
                const float sm = div_const<IEEE>(ssum[pp], 3.0f, r3), lm = div_const<IEEE>(lsum[pp], 3.0f, r3);   // .mean(1, True)
                const float l = xadd(xmul(0.85f, sm), xmul(0.15f, lm));               // losses.py:115
                if (p.loss_map) p.loss_map[((long long)b * H + y) * W + x] = l;
                acc += l;
            }
        }
        if (p.partial) {
            const float tot = block_sum(acc, red);
            if (threadIdx.x == 0) p.partial[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
        }
    }
}

template <int MODE, int CK, int TH, int TW, int NT, bool IL>
__global__ void __launch_bounds__(NT) warp_photo_fwd_kernel(const __grid_constant__ WPParams p)
{
    constexpr int RH = TH + 2, RP = TW + 2;
    static_assert(TH * TW % NT == 0 && NT % TW == 0 && (NT / TW) * (TH * TW / NT) == TH, "tile / thread mapping");
    __shared__ float sx[CK * RH * RP];
    __shared__ float sy[CK * RH * RP];
    __shared__ float cam[24];
    __shared__ float red[NT / 32];

    const int planes_per_b = (MODE == MODE_DIRECT) ? p.C / CK : 1;
    const int b = blockIdx.z / planes_per_b, ch0 = (blockIdx.z % planes_per_b) * CK;
    const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;

    if (MODE == MODE_WARP) {
        stage_camera(p, b, b, cam);
        __syncthreads();
    }
    const bool bad = fill_region<MODE, CK, RH, RP, 1, TH, TW, IL>(p, cam, b, ch0, ty0, tx0, sx, sy, true);
    const int slow = __syncthreads_or(bad ? 1 : 0) | !p.div_exact;
    reflect_fixup<CK, RH, RP>(sx, ty0 - 1, tx0 - 1, p.H, p.W);
    reflect_fixup<CK, RH, RP>(sy, ty0 - 1, tx0 - 1, p.H, p.W);
    __syncthreads();
    if (slow) fwd_phase2<true, CK, TH, TW, NT>(p, sx, sy, red, b, ch0, ty0, tx0);
    else fwd_phase2<false, CK, TH, TW, NT>(p, sx, sy, red, b, ch0, ty0, tx0);
}

// Deterministic final reduction of per-CTA partial sums: out[0] = sum(partial[0..n)) * scale.
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float *partial, long long n, double scale, float *out)
{
    __shared__ double sh[32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partial[i];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = sh[threadIdx.x];
        acc = warp_sum_d(acc);
        if (threadIdx.x == 0) out[0] = (float)(acc * scale);
    }
}

// grad_P[b][e] = sum over the CTAs of batch element b of partial[cta][e]  (fixed order -> deterministic)
__global__ void __launch_bounds__(256) reduce_gP_kernel(const float *partial, int ctas_per_b, float *gP, const float *skip_flag)
{
    __shared__ double sh[8];
    if (skip_flag && __ldg(skip_flag) != 0.0f) return;      // conditional backward: the partials were not produced
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

// Both reductions of the single-sweep path in ONE launch (a single pair is launch-latency bound: config C1): blocks 0 .. 12 B - 1
// reduce grad_P as reduce_gP_kernel does, the last block reduces the loss partials.
__global__ void __launch_bounds__(256) reduce_loss_gP_kernel(const float *loss_partial, long long n, double scale, float *loss,
                                                             const float *gp_partial, int ctas_per_b, int B, float *gP, const float *skip_flag)
{
    __shared__ double sh[8];
    const bool loss_block = (int)blockIdx.x == 12 * B;
    if (!loss_block && skip_flag && __ldg(skip_flag) != 0.0f) return;
    double acc = 0.0;
    const int b = blockIdx.x / 12, e = blockIdx.x % 12;
    if (loss_block) {
        for (long long i = threadIdx.x; i < n; i += 256) acc += (double)loss_partial[i];
    } else {
        for (int i = threadIdx.x; i < ctas_per_b; i += 256) acc += (double)gp_partial[((long long)b * ctas_per_b + i) * 12 + e];
    }
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += sh[i];
        if (loss_block) loss[0] = (float)(t * scale);
        else gP[b * 12 + e] = (float)t;
    }
}

int launch_reduce_loss_gP(const float *loss_partial, long long n, double scale, float *loss, const float *gp_partial, int ctas_per_b, int B,
                          float *gP, cudaStream_t st, const float *skip_flag)
{
    reduce_loss_gP_kernel<<<B * 12 + 1, 256, 0, st>>>(loss_partial, n, scale, loss, gp_partial, ctas_per_b, B, gP, skip_flag);
    count_launch();
    return finish_launch("reduce_loss_gP_kernel");
}

int launch_reduce_partials(const float *partial, long long n, double scale, float *out, cudaStream_t st)
{
    reduce_partials_kernel<<<1, 1024, 0, st>>>(partial, n, scale, out);
    count_launch();
    return finish_launch("reduce_partials_kernel");
}

int launch_reduce_gP(const float *partial, int ctas_per_b, int B, float *gP, cudaStream_t st, const float *skip_flag)
{
    reduce_gP_kernel<<<B * 12, 256, 0, st>>>(partial, ctas_per_b, gP, skip_flag);
    count_launch();
    return finish_launch("reduce_gP_kernel");
}

// ================================================================================================
// Backward kernel
// ================================================================================================
template <bool IEEE, int CK, int TH, int TW, int NT, bool NEED_GY>
__device__ __forceinline__ void bwd_phaseB(const WPParams &p, const float *sx, const float *sy, float *G,
                                           int b, int ch0, int ty0, int tx0, float gscal)
{
    constexpr int RH2 = TH + 4, RP2 = TW + 4, RH1 = TH + 2, RP1 = TW + 2;
    constexpr int NG = NEED_GY ? 4 : 3, NPB = RH1 / 3;
    const int H = p.H, W = p.W;
    const float invC = 1.0f / (float)p.C;
    for (int task = threadIdx.x; task < 3 * RP1; task += NT) {
        const int cg = task / RP1, cc = task - cg * RP1;      // strip cg covers centre rows cg*NPB .. +NPB-1
        const int cx = tx0 - 1 + cc;
#pragma unroll
        for (int ch = 0; ch < CK; ch++) {
            float S[NPB][5];
            const float *wx = sx + (ch * RH2 + cg * NPB) * RP2 + cc;
            const float *wy = sy + (ch * RH2 + cg * NPB) * RP2 + cc;
            window_sums<NPB, RP2>(wx, wy, S);
#pragma unroll
            for (int pp = 0; pp < NPB; pp++) {
                const int cr = cg * NPB + pp, cy = ty0 - 1 + cr;
                float ga = 0.f, gb = 0.f, gc = 0.f, gay = 0.f;
                if (cy >= 0 && cy < H && cx >= 0 && cx < W) {
                    SsimVals v;
                    ssim_finish<IEEE>(S[pp], v);
                    float gl = gscal;
                    if (p.g_loss_map) gl = __ldg(p.g_loss_map + ((long long)b * H + cy) * W + cx);
                    float gs = 0.85f * invC * gl;
                    if (p.g_ssim) gs += __ldg(p.g_ssim + (((long long)b * p.C + ch0 + ch) * H + cy) * W + cx);
                    if (!(v.sraw >= 0.0f && v.sraw <= 1.0f)) gs = 0.0f;       // clamp passes gradient on [0,1]
                    const float rdn = __frcp_rn(v.dn);
                    const float dA = v.A2 - v.A1, dB = v.B2 - v.B1;
                    const float h = (-0.5f / 9.0f) * gs * rdn;                 // d ssim/dQ = -1/2, box filter 1/9
                    const float tq = 2.0f * v.Q;
                    ga = h * (2.0f * v.muy * dA - tq * v.mux * dB);
                    gb = -h * v.Q * v.B1;
                    gc = h * 2.0f * v.A1;
                    if (NEED_GY) gay = h * (2.0f * v.mux * dA - tq * v.muy * dB);
                }
                float *g = G + ((ch * NG) * RH1 + cr) * RP1 + cc;
                g[0] = ga;
                g[RH1 * RP1] = gb;
                g[2 * RH1 * RP1] = gc;
                if (NEED_GY) g[3 * RH1 * RP1] = gay;
            }
        }
    }
}

template <int MODE, int CK, int TH, int TW, int NT, bool NEED_GY, bool IL>
__global__ void __launch_bounds__(NT) warp_photo_bwd_kernel(const __grid_constant__ WPParams p)
{
    constexpr int RH2 = TH + 4, RP2 = TW + 4;      // x / y region (2-pixel halo)
    constexpr int RH1 = TH + 2, RP1 = TW + 2;      // centre region (1-pixel halo)
    constexpr int NG = NEED_GY ? 4 : 3;            // coefficient planes per channel
    constexpr int NPC = TH * TW / NT;              // owners per thread in phase C
    static_assert(RH1 % 3 == 0 && TH * TW % NT == 0 && NT % TW == 0 && (NT / TW) * NPC == TH, "tile / thread mapping");
    extern __shared__ float smem[];
    float *sx = smem;                              // [CK][RH2][RP2]
    float *sy = sx + CK * RH2 * RP2;
    float *G = sy + CK * RH2 * RP2;                // [CK][NG][RH1][RP1]
    float *cam = G + CK * NG * RH1 * RP1;          // 24
    float *red = cam + 24;                         // NT/32 * 12

    const int planes_per_b = (MODE == MODE_DIRECT) ? p.C / CK : 1;
    const int b = blockIdx.z / planes_per_b, ch0 = (blockIdx.z % planes_per_b) * CK;
    const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
    const int H = p.H, W = p.W;
    const float invC = 1.0f / (float)p.C;
    // uniform upstream gradient when no per-pixel map is given
    float gscal = 0.0f;
    if (!p.g_loss_map && !p.g_ssim) gscal = (p.g_scalar ? __ldg(p.g_scalar) : 1.0f) * p.g_scale;

    if (MODE == MODE_WARP) {
        stage_camera(p, b, b, cam);
        __syncthreads();
    }
    // ---- phase A ---------------------------------------------------------------------------
    const bool bad = fill_region<MODE, CK, RH2, RP2, 2, TH, TW, IL>(p, cam, b, ch0, ty0, tx0, sx, sy, false);
    const int slow = __syncthreads_or(bad ? 1 : 0) | !p.div_exact;
    reflect_fixup<CK, RH2, RP2>(sx, ty0 - 2, tx0 - 2, H, W);
    reflect_fixup<CK, RH2, RP2>(sy, ty0 - 2, tx0 - 2, H, W);
    __syncthreads();

    // ---- phase B: coefficient planes at every centre of the 1-pixel halo ----------------------
    if (slow) bwd_phaseB<true, CK, TH, TW, NT, NEED_GY>(p, sx, sy, G, b, ch0, ty0, tx0, gscal);
    else bwd_phaseB<false, CK, TH, TW, NT, NEED_GY>(p, sx, sy, G, b, ch0, ty0, tx0, gscal);
    __syncthreads();

    // ---- phase C: owners ---------------------------------------------------------------------
    const int col = threadIdx.x % TW, rg = threadIdx.x / TW;
    const int x = tx0 + col;
    // column multiplicities of the folded reflect padding (centre col x+dx contributes wcol[dx+1] times)
    float wcol[3];
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) {
        const int cx = x + dx;
        float w = (cx >= 0 && cx < W) ? 1.0f : 0.0f;
        if (dx != 0 && ((cx == 0 && x == 1) || (cx == W - 1 && x == W - 2))) w += 1.0f;
        wcol[dx + 1] = w;
    }
    float gP[12];
#pragma unroll
    for (int e = 0; e < 12; e++) gP[e] = 0.f;
    const Img32 src = cta_image(p.src, b);
    const PixConst kc = pix_const(p);
    const bool use_mask = p.use_mask != 0;
    float *gsrc_b = p.g_src.p ? p.g_src.p + (long long)b * p.g_src.sb : nullptr;
    const int gs_sc = (int)p.g_src.sc, gs_sh = (int)p.g_src.sh, gs_sw = (int)p.g_src.sw;

#pragma unroll
    for (int pp = 0; pp < NPC; pp++) {
        const int row = rg * NPC + pp, y = ty0 + row;
        const bool inside = (y < H && x < W);
        float wrow[3];
#pragma unroll
        for (int dy = -1; dy <= 1; dy++) {
            const int cy = y + dy;
            float w = (cy >= 0 && cy < H) ? 1.0f : 0.0f;
            if (dy != 0 && ((cy == 0 && y == 1) || (cy == H - 1 && y == H - 2))) w += 1.0f;
            wrow[dy + 1] = w;
        }
        float gl = gscal;
        if (p.g_loss_map && inside) gl = __ldg(p.g_loss_map + ((long long)b * H + y) * W + x);
        const float gl1 = 0.15f * invC * gl;
        float gsyn[CK];
#pragma unroll
        for (int ch = 0; ch < CK; ch++) {
            float acc[NG];
#pragma unroll
            for (int k = 0; k < NG; k++) acc[k] = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
#pragma unroll
                for (int k = 0; k < NG; k++) {
                    const float *g = G + ((ch * NG + k) * RH1 + row + dy) * RP1 + col;
                    const float rs = wcol[0] * g[0] + wcol[1] * g[1] + wcol[2] * g[2];
                    acc[k] += wrow[dy] * rs;
                }
            }
            const float xj = sx[(ch * RH2 + row + 2) * RP2 + col + 2], yj = sy[(ch * RH2 + row + 2) * RP2 + col + 2];
            const float df = xj - yj;
            const float sg = (df > 0.f) ? 1.f : ((df < 0.f) ? -1.f : 0.f);
            const float gxj = acc[0] + 2.0f * xj * acc[1] + yj * acc[2] + gl1 * sg;
            gsyn[ch] = gxj;
            if (MODE == MODE_DIRECT && inside) {
                const long long o = (((long long)b * p.C + ch0 + ch) * H + y) * W + x;
                if (p.g_x) p.g_x[o] = gxj;
                if (NEED_GY && p.g_y) p.g_y[o] = acc[NG - 1] + 2.0f * yj * acc[1] + xj * acc[2] - gl1 * sg;
            }
        }
        if (MODE == MODE_WARP && inside) {
            const int pixo = y * W + x;
            const long long pixi = (long long)b * H * W + pixo;
            const float d = __ldg(p.depth + pixi);
            Proj pr;
            project_pixel(cam, kc, x, y, d, pr);
            Samp s;
            sampler_setup(kc, pr.gx, pr.gy, s);
            int o[4];
            tap_offsets(src, s, o);
            const int go0 = s.y0 * gs_sh + s.x0 * gs_sw;
            float gix = 0.f, giy = 0.f;
            const float wxx = s.ix - floorf(s.ix), ex = 1.0f - wxx;
            const float wyy = s.iy - floorf(s.iy), ey = 1.0f - wyy;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const float gsy = use_mask ? gsyn[ch] * pr.valid : gsyn[ch];
                float v[4];
                gather_taps<IL>(src, s, o, ch, v);
                gix += gsy * ((v[1] - v[0]) * ey + (v[3] - v[2]) * wyy);
                giy += gsy * ((v[2] - v[0]) * ex + (v[3] - v[1]) * wxx);
                if (gsrc_b) {
                    float *gp = gsrc_b + go0 + ch * gs_sc;
                    if (s.in00) atomicAdd(gp, gsy * s.nw);
                    if (s.in01) atomicAdd(gp + gs_sw, gsy * s.ne);
                    if (s.in10) atomicAdd(gp + gs_sh, gsy * s.sw);
                    if (s.in11) atomicAdd(gp + gs_sh + gs_sw, gsy * s.se);
                }
            }
            // sample position -> pixel coordinate -> camera point        
            const float gu = gix * s.mx * (kc.half_w * 2.0f / kc.wm1);
            const float gv = giy * s.my * (kc.half_h * 2.0f / kc.hm1);
            const float rz = __frcp_rn(pr.z);
            const float gc0 = gu * rz, gc1 = gv * rz;
            const float gc2 = -(gu * pr.c0 + gv * pr.c1) * rz * rz;
            const float *P = cam + 9;
            const float q0 = P[0] * pr.r0 + P[1] * pr.r1 + P[2] * pr.r2;
            const float q1 = P[4] * pr.r0 + P[5] * pr.r1 + P[6] * pr.r2;
            const float q2 = P[8] * pr.r0 + P[9] * pr.r1 + P[10] * pr.r2;
            // d u/d depth = (q0*tz - t0*q2)/z^2 with c = depth*q + t: the well-conditioned form of
            // gc . q (which cancels ~100x in fp32); agrees with exact arithmetic to ~4e-7.
            const float tz = P[11] + kc.eps;
            const float du = q0 * tz - P[3] * q2, dv = q1 * tz - P[7] * q2;
            p.g_depth[pixi] = (gu * du + gv * dv) * rz * rz;
            gP[0] += gc0 * pr.X0; gP[1] += gc0 * pr.X1; gP[2] += gc0 * pr.X2; gP[3] += gc0;
            gP[4] += gc1 * pr.X0; gP[5] += gc1 * pr.X1; gP[6] += gc1 * pr.X2; gP[7] += gc1;
            gP[8] += gc2 * pr.X0; gP[9] += gc2 * pr.X1; gP[10] += gc2 * pr.X2; gP[11] += gc2;
        }
    }

    if (MODE == MODE_WARP && p.gP_partial) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int e = 0; e < 12; e++) {
            const float v = warp_sum(gP[e]);
            if (lane == 0) red[wid * 12 + e] = v;
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            float t = 0.f;
            for (int w = 0; w < NT / 32; w++) t += red[w * 12 + threadIdx.x];
            const long long cta = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            p.gP_partial[cta * 12 + threadIdx.x] = t;
        }
    }
}

// ================================================================================================
// Host launchers
// ================================================================================================
constexpr int F_TH = 16, F_TW = 64, F_NT = 256;
constexpr int B_TH = 16, B_TW = 64, B_NT = 256;

static inline dim3 tile_grid(int B, int H, int W, int TH, int TW)
{
    return dim3((W + TW - 1) / TW, (H + TH - 1) / TH, B);
}

static size_t partial_count(int B, int H, int W)
{
    const dim3 g = tile_grid(B, H, W, F_TH < B_TH ? F_TH : B_TH, F_TW < B_TW ? F_TW : B_TW);
    return (size_t)g.x * g.y * g.z;
}

template <int CK, bool NEED_GY>
static constexpr size_t bwd_smem_bytes()
{
    return sizeof(float) * (2 * CK * (B_TH + 4) * (B_TW + 4) + CK * (NEED_GY ? 4 : 3) * (B_TH + 2) * (B_TW + 2) + 24 + (B_NT / 32) * 12);
}

template <int MODE, int CK, bool NEED_GY, bool IL>
static int launch_bwd(const WPParams &p, dim3 grid, cudaStream_t st)
{
    auto kern = warp_photo_bwd_kernel<MODE, CK, B_TH, B_TW, B_NT, NEED_GY, IL>;
    constexpr size_t smem = bwd_smem_bytes<CK, NEED_GY>();
    static bool configured_dev[64] = {};          // per device (and per template instance): one process may drive several GPUs
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    bool &configured = configured_dev[dev_id & 63];
    if (!configured) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    kern<<<grid, B_NT, smem, st>>>(p);
    count_launch();
    return finish_launch("warp_photo_bwd_kernel");
}

bool view_fits_int32(const ImgView &v, int C, int H, int W)
{
    const long long ext = llabs(v.sc) * (C - 1) + llabs(v.sh) * (H - 1) + llabs(v.sw) * (W - 1);
    return ext < (1ll << 30) && llabs(v.sc) < (1ll << 30) && llabs(v.sh) < (1ll << 30) && llabs(v.sw) < (1ll << 30);
}

int fill_common(WPParams &p, int B, int C, int H, int W, int padding_mode, int use_mask, float eps, cudaStream_t st)
{
    E2E_REQUIRE(B > 0 && H >= 2 && W >= 2, "B=%d H=%d W=%d: need B>0, H>=2, W>=2 (reflection padding)", B, H, W);
    E2E_REQUIRE(padding_mode == 0 || padding_mode == 1, "padding_mode must be 0 (zeros) or 1 (border)");
    E2E_REQUIRE((long long)B * C <= 65535, "B*C exceeds gridDim.z");
    E2E_REQUIRE((long long)H * W * 3 < (1ll << 30), "image too large for 32-bit in-image offsets");
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.border = padding_mode; p.use_mask = use_mask; p.eps = eps;
    p.wm1 = (float)(W - 1); p.hm1 = (float)(H - 1);
    p.half_w = (float)W / 2; p.half_h = (float)H / 2;
    const DivC dW = host_divc(p.wm1, st), dH = host_divc(p.hm1, st), d9 = host_divc(9.0f, st), d3 = host_divc(3.0f, st);
    p.rcpW = dW.rcp; p.rcpH = dH.rcp;
    p.div_exact = dW.exact && dH.exact && d9.exact && d3.exact;
    return 0;
}

int set_views(WPParams &p, const float *a, const int64_t as[4], const float *b, const int64_t bs[4], int C)
{
    p.src = make_view(a, as);
    p.tgt = make_view(b, bs);
    E2E_REQUIRE(view_fits_int32(p.src, C, p.H, p.W) && view_fits_int32(p.tgt, C, p.H, p.W),
                "image strides do not fit 32-bit in-image offsets");
    return 0;
}

}  // namespace e2e

using namespace e2e;

extern "C" {

size_t e2e_warp_photo_workspace_bytes(int B, int H, int W)
{
    // forward: one float per CTA; backward: the streaming kernel's partials (or 12 floats per CTA of the tile kernel)
    const size_t tile = partial_count(B, H, W) * 12 * sizeof(float) + 256, stream = stream_workspace_bytes(B, H, W);
    return tile > stream ? tile : stream;
}

int e2e_warp_photo_fwd(const float *depth, const float *inv_K, const float *K, const float *T,
                       const float *src, const int64_t src_strides[4], const float *tgt, const int64_t tgt_strides[4],
                       int B, int H, int W, int padding_mode, int use_mask, float eps,
                       float *syn, float *valid, float *pix, float *loss_map, float *loss_mean,
                       void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    WPParams p = {};
    E2E_REQUIRE(depth && inv_K && K && T && src && tgt, "null input pointer");
    if (int rc = fill_common(p, B, 3, H, W, padding_mode, use_mask, eps, st)) return rc;
    p.depth = depth; p.inv_K = inv_K; p.K = K; p.T = T;
    if (int rc = set_views(p, src, src_strides, tgt, tgt_strides, 3)) return rc;
    p.syn = syn; p.valid = valid; p.pix = pix; p.loss_map = loss_map;
    // Default: the value-only instances of the streaming kernel (csrc/warp_photo_fused.cu); E2E_FWD_TILE=1 selects the tile kernel.
    static const bool use_tile = [] { const char *e = getenv("E2E_FWD_TILE"); return e && e[0] == '1'; }();
    if (!use_tile && H <= 8189 && W <= 8189 && workspace && workspace_bytes >= stream_workspace_bytes(B, H, W))
        return launch_stream(p, B, H, W, loss_mean, nullptr, workspace, workspace_bytes, st);
    const dim3 grid = tile_grid(B, H, W, F_TH, F_TW);
    const size_t nct = (size_t)grid.x * grid.y * grid.z;
    if (loss_mean) {
        E2E_REQUIRE(workspace && workspace_bytes >= nct * sizeof(float), "workspace too small for loss_mean");
        p.partial = (float *)workspace;
    }
    const bool il = (p.src.sc == 1 && p.tgt.sc == 1);     // interleaved RGB (channels-last memory)
    if (il) warp_photo_fwd_kernel<MODE_WARP, 3, F_TH, F_TW, F_NT, true><<<grid, F_NT, 0, st>>>(p);
    else warp_photo_fwd_kernel<MODE_WARP, 3, F_TH, F_TW, F_NT, false><<<grid, F_NT, 0, st>>>(p);
    count_launch();
    if (int rc = finish_launch("warp_photo_fwd_kernel")) return rc;
    if (loss_mean) {
        reduce_partials_kernel<<<1, 1024, 0, st>>>(p.partial, (long long)nct, 1.0 / ((double)B * H * W), loss_mean);
        count_launch();
        if (int rc = finish_launch("reduce_partials_kernel")) return rc;
    }
    return 0;
}

int e2e_warp_photo_bwd(const float *depth, const float *inv_K, const float *K, const float *T,
                       const float *src, const int64_t src_strides[4], const float *tgt, const int64_t tgt_strides[4],
                       int B, int H, int W, int padding_mode, int use_mask, float eps,
                       const float *grad_loss_map, const float *grad_scalar, float scalar_scale,
                       float *grad_depth, float *grad_src, const int64_t grad_src_strides[4], float *grad_P,
                       void *workspace, size_t workspace_bytes, void *stream)
{
    return e2e_warp_photo_bwd_cond(depth, inv_K, K, T, src, src_strides, tgt, tgt_strides, B, H, W, padding_mode, use_mask, eps,
                                   grad_loss_map, grad_scalar, scalar_scale, nullptr, grad_depth, grad_src, grad_src_strides, grad_P,
                                   workspace, workspace_bytes, stream);
}

int e2e_warp_photo_bwd_cond(const float *depth, const float *inv_K, const float *K, const float *T,
                            const float *src, const int64_t src_strides[4], const float *tgt, const int64_t tgt_strides[4],
                            int B, int H, int W, int padding_mode, int use_mask, float eps,
                            const float *grad_loss_map, const float *grad_scalar, float scalar_scale, const float *skip_if_nonzero,
                            float *grad_depth, float *grad_src, const int64_t grad_src_strides[4], float *grad_P,
                            void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    WPParams p = {};
    E2E_REQUIRE(depth && inv_K && K && T && src && tgt && grad_depth, "null pointer");
    if (int rc = fill_common(p, B, 3, H, W, padding_mode, use_mask, eps, st)) return rc;
    p.depth = depth; p.inv_K = inv_K; p.K = K; p.T = T;
    if (int rc = set_views(p, src, src_strides, tgt, tgt_strides, 3)) return rc;
    p.g_loss_map = grad_loss_map; p.g_scalar = grad_scalar; p.g_scale = scalar_scale;
    p.skip_flag = skip_if_nonzero;
    E2E_REQUIRE(!skip_if_nonzero || grad_loss_map, "the conditional backward takes the upstream gradient as a map");
    p.g_depth = grad_depth;
    if (grad_src) {
        E2E_REQUIRE(grad_src_strides, "grad_src needs strides");
        p.g_src = make_view_w(grad_src, grad_src_strides);
        const ImgView gv{grad_src, p.g_src.sb, p.g_src.sc, p.g_src.sh, p.g_src.sw};
        E2E_REQUIRE(view_fits_int32(gv, 3, H, W), "grad_src strides do not fit 32-bit in-image offsets");
    }
    // Default: the streaming kernel (csrc/warp_photo_fused.cu) with the upstream gradient read per pixel -- it re-derives the
    // forward quantities in the same sweep, half the time of the tile kernel below.  E2E_BWD_TILE=1 selects the tile kernel.
    static const bool use_tile = [] { const char *e = getenv("E2E_BWD_TILE"); return e && e[0] == '1'; }();
    if ((!use_tile || skip_if_nonzero) && H <= 8189 && W <= 8189)
        return launch_stream(p, B, H, W, nullptr, grad_P, workspace, workspace_bytes, st);
    E2E_REQUIRE(!skip_if_nonzero, "the conditional backward needs H, W <= 8189");
    const dim3 grid = tile_grid(B, H, W, B_TH, B_TW);
    const size_t nct = (size_t)grid.x * grid.y * grid.z;
    if (grad_P) {
        E2E_REQUIRE(workspace && workspace_bytes >= nct * 12 * sizeof(float), "workspace too small for grad_P");
        p.gP_partial = (float *)workspace;
    }
    const bool il = (p.src.sc == 1 && p.tgt.sc == 1);
    if (int rc = il ? launch_bwd<MODE_WARP, 3, false, true>(p, grid, st) : launch_bwd<MODE_WARP, 3, false, false>(p, grid, st)) return rc;
    if (grad_P) {
        reduce_gP_kernel<<<B * 12, 256, 0, st>>>(p.gP_partial, (int)(grid.x * grid.y), grad_P, nullptr);
        count_launch();
        if (int rc = finish_launch("reduce_gP_kernel")) return rc;
    }
    return 0;
}

int e2e_ssim_fwd(const float *x, const int64_t x_strides[4], const float *y, const int64_t y_strides[4],
                 int B, int C, int H, int W, float *ssim_map, float *loss_map, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    WPParams p = {};
    E2E_REQUIRE(x && y && C >= 1, "null input / bad C");
    E2E_REQUIRE(!loss_map || C == 3, "loss_map output requires C == 3 (photometric_loss, losses.py:97-117)");
    if (int rc = fill_common(p, B, C, H, W, 1, 0, 0.f, st)) return rc;
    if (int rc = set_views(p, x, x_strides, y, y_strides, C)) return rc;
    p.ssim = ssim_map; p.loss_map = loss_map;
    static const bool use_tile = [] { const char *e = getenv("E2E_FWD_TILE"); return e && e[0] == '1'; }();
    if (C == 3 && !use_tile && H <= 8189 && W <= 8189 && (ssim_map || loss_map)) {      // the streaming kernel, value only
        const ImgView vx = p.src, vy = p.tgt;
        if (view_fits_int32(vx, 3, H, W) && view_fits_int32(vy, 3, H, W)) return launch_ssim_stream_bwd(p, B, H, W, st);
    }
    dim3 grid = tile_grid(B, H, W, F_TH, F_TW);
    if (C == 3) {
        warp_photo_fwd_kernel<MODE_DIRECT, 3, F_TH, F_TW, F_NT, false><<<grid, F_NT, 0, st>>>(p);
    } else {
        grid.z = B * C;
        warp_photo_fwd_kernel<MODE_DIRECT, 1, F_TH, F_TW, F_NT, false><<<grid, F_NT, 0, st>>>(p);
    }
    count_launch();
    return finish_launch("ssim_fwd_kernel");
}

int e2e_ssim_bwd(const float *x, const int64_t x_strides[4], const float *y, const int64_t y_strides[4],
                 int B, int C, int H, int W, const float *grad_ssim, const float *grad_loss_map,
                 float *grad_x, float *grad_y, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    WPParams p = {};
    E2E_REQUIRE(x && y && C >= 1 && (grad_ssim || grad_loss_map), "null input / no upstream gradient");
    E2E_REQUIRE(!grad_loss_map || C == 3, "grad_loss_map requires C == 3");
    if (int rc = fill_common(p, B, C, H, W, 1, 0, 0.f, st)) return rc;
    if (int rc = set_views(p, x, x_strides, y, y_strides, C)) return rc;
    p.g_ssim = grad_ssim; p.g_loss_map = grad_loss_map;
    p.g_x = grad_x; p.g_y = grad_y;
    // C == 3 and only d / d x wanted (the reference never differentiates the target): the streaming kernel
    static const bool use_tile = [] { const char *e = getenv("E2E_BWD_TILE"); return e && e[0] == '1'; }();
    if (C == 3 && grad_x && !grad_y && !use_tile && !(grad_ssim && grad_loss_map) && H <= 8189 && W <= 8189) {
        const ImgView vx = p.src, vy = p.tgt;
        if (view_fits_int32(vx, 3, H, W) && view_fits_int32(vy, 3, H, W)) return launch_ssim_stream_bwd(p, B, H, W, st);
    }
    dim3 grid = tile_grid(B, H, W, B_TH, B_TW);
    if (C == 3) return launch_bwd<MODE_DIRECT, 3, true, false>(p, grid, st);
    grid.z = B * C;
    return launch_bwd<MODE_DIRECT, 1, true, false>(p, grid, st);
}

}  // extern "C"
