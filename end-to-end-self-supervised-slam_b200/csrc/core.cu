// Library bookkeeping: error text, launch counter, constant-divisor verification.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>

#include "common.cuh"

namespace e2e {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int finish_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// One thread per mantissa: compares the 3-instruction quotient with IEEE division for x in [1,2)
// and [2^20, 2^21) (the sequence is scale invariant away from under/overflow).
__global__ void verify_divisor_kernel(float d, float rcp, unsigned int *bad)
{
    const unsigned int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (1u << 23)) return;
    unsigned int nbad = 0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        float x = __uint_as_float(0x3f800000u | m);
        if (s) x = x * 1048576.0f;
        const float q = __fmul_rn(x, rcp);
        const float r = __fmaf_rn(-d, q, x);
        const float q2 = __fmaf_rn(r, rcp, q);
        nbad += (__float_as_uint(q2) != __float_as_uint(__fdiv_rn(x, d)));
        // negative divisor/operand symmetry: sign handling is exact in all three instructions
    }
    if (nbad) atomicAdd(bad, nbad);
}

static std::mutex g_div_mutex;
static std::map<uint32_t, int> g_div_cache;   // float bits -> exact?

static int prepare_divisor(float d, cudaStream_t stream)
{
    if (!(d == d) || d == 0.f || isinf(d)) return 0;
    uint32_t bits;
    memcpy(&bits, &d, 4);
    {
        std::lock_guard<std::mutex> lk(g_div_mutex);
        auto it = g_div_cache.find(bits);
        if (it != g_div_cache.end()) return it->second;
    }
    unsigned int *bad = nullptr;
    if (cudaMalloc(&bad, sizeof(unsigned int)) != cudaSuccess) return -1;
    cudaMemsetAsync(bad, 0, sizeof(unsigned int), stream);
    verify_divisor_kernel<<<(1u << 23) / 256, 256, 0, stream>>>(d, 1.0f / d, bad);
    count_launch();
    unsigned int h = 1;
    cudaMemcpyAsync(&h, bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
    const cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(bad);
    if (e != cudaSuccess) {
        set_error("e2e_prepare_divisor: %s", cudaGetErrorString(e));
        return -1;
    }
    const int exact = (h == 0) ? 1 : 0;
    std::lock_guard<std::mutex> lk(g_div_mutex);
    g_div_cache[bits] = exact;
    return exact;
}

DivC host_divc(float d, cudaStream_t stream)
{
    DivC k;
    k.d = d;
    k.rcp = 1.0f / d;
    const int ex = prepare_divisor(d, stream);
    k.exact = ex > 0 ? 1 : 0;
    return k;
}

}  // namespace e2e

namespace e2e {

// 8-bit frames -> the reference's float frames: colors /= 255.0 (train_depth.py:255, online_adaption.py:215) done on the
// device, correctly rounded division like the host's, so only a quarter of the bytes cross PCIe.
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const uchar4 *in, long long n4, float4 *out)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const uchar4 v = in[i];
        out[i] = make_float4(__fdiv_rn((float)v.x, 255.0f), __fdiv_rn((float)v.y, 255.0f), __fdiv_rn((float)v.z, 255.0f),
                             __fdiv_rn((float)v.w, 255.0f));
    }
}
// scalar form: the tail of an aligned array, or the whole of a misaligned one (a contiguous uint8 VIEW may start at any byte)
__global__ void __launch_bounds__(256) u8_to_unit_scalar_kernel(const unsigned char *in, long long from, long long n, float *out)
{
    for (long long i = from + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        out[i] = __fdiv_rn((float)in[i], 255.0f);
}

}  // namespace e2e

extern "C" {

int e2e_u8_to_unit(const unsigned char *in, long long n, float *out, void *stream)
{
    E2E_REQUIRE(in && out && n > 0, "u8_to_unit: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (((uintptr_t)in) & 3) == 0 && (((uintptr_t)out) & 15) == 0;
    const long long n4 = vec ? n / 4 : 0;
    if (n4 > 0) {
        long long blocks = (n4 + 255) / 256;
        if (blocks > e2e::kNumSMs * 16) blocks = e2e::kNumSMs * 16;
        e2e::u8_to_unit_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uchar4 *)in, n4, (float4 *)out);
        e2e::count_launch();
    }
    if (n4 * 4 < n) {
        long long blocks = (n - n4 * 4 + 255) / 256;
        if (blocks > e2e::kNumSMs * 16) blocks = e2e::kNumSMs * 16;
        e2e::u8_to_unit_scalar_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, n4 * 4, n, out);
        e2e::count_launch();
    }
    return e2e::finish_launch("u8_to_unit");
}

int e2e_abi_version(void) { return 1; }
const char *e2e_last_error(void) { return e2e::g_err; }
unsigned long long e2e_launch_count(void) { return e2e::g_launches.load(); }
int e2e_prepare_divisor(float d, void *stream) { return e2e::prepare_divisor(d, (cudaStream_t)stream); }

}  // extern "C"
