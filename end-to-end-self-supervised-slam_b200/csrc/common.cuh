// Shared helpers for the sm_100a kernels: exact-order fp32 primitives, constant-divisor division,
// strided image views, block reductions, launch bookkeeping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/e2e_slam_b200.h"

namespace e2e {

// ---------------------------------------------------------------------------------------------
// Exact-order arithmetic.  The forward path reproduces the reference's CPU results bit for bit
// (oracle/warp_photo_oracle.c is the operation-order specification).  The _rn intrinsics are
// never contracted into FMAs by nvcc, so each call below is exactly one IEEE rounding.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }

// x / d for a divisor d known before launch:  q = x*c; r = fma(-d, q, x); q' = fma(r, c, q) with
// c = RN(1/d).  Correctly rounded for every normal x away from the under/overflow range iff d passes
// the exhaustive mantissa check of e2e_prepare_divisor(); `exact` carries that verdict.
struct DivC {
    float d;
    float rcp;
    int exact;
};

__device__ __forceinline__ float xdivc(float x, const DivC k)
{
    const uint32_t e = (__float_as_uint(x) >> 23) & 0xffu;          // biased exponent
    if (k.exact && (e - 32u) < 190u) {                             // 2^-95 <= |x| < 2^95
        const float q = __fmul_rn(x, k.rcp);
        const float r = __fmaf_rn(-k.d, q, x);
        return __fmaf_rn(r, k.rcp, q);
    }
    return __fdiv_rn(x, k.d);                                      // zeros, denormals, inf/nan, huge
}

// ---------------------------------------------------------------------------------------------
// Strided 4-D image view (element strides): consumes the reference's NCHW views of NHWC memory.
// ---------------------------------------------------------------------------------------------
struct ImgView {
    const float *p;
    long long sb, sc, sh, sw;
};

struct ImgViewW {
    float *p;
    long long sb, sc, sh, sw;
};

inline ImgView make_view(const float *p, const int64_t s[4]) { return ImgView{p, s[0], s[1], s[2], s[3]}; }
inline ImgViewW make_view_w(float *p, const int64_t s[4]) { return ImgViewW{p, s[0], s[1], s[2], s[3]}; }

// ---------------------------------------------------------------------------------------------
// Reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in thread 0.  `scratch` needs blockDim.x/32 floats.
__device__ __forceinline__ float block_sum(float v, float *scratch)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    float r = 0.f;
    if (wid == 0) {
        r = lane < nw ? scratch[lane] : 0.f;
        r = warp_sum(r);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// Host side bookkeeping
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int finish_launch(const char *what);   // cudaGetLastError -> return code, records text
DivC host_divc(float d, cudaStream_t stream);

constexpr int kNumSMs = 148;   // B200

#define E2E_REQUIRE(cond, ...)               \
    do {                                     \
        if (!(cond)) {                       \
            ::e2e::set_error(__VA_ARGS__);   \
            return E2E_ERR_BAD_ARG;          \
        }                                    \
    } while (0)

}  // namespace e2e
