// Micro-benchmark: issue throughput of scalar vs packed (f32x2) fp32 add / mul / fma on sm_100a,
// and LDS.32 / LDS.64 / LDS.128 shared-memory load throughput.  Prints warp-instructions per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int OP>
__global__ void __launch_bounds__(1024) k_scalar(float *out, float a, float b)
{
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b));
            if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(b), "f"(a));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
__global__ void __launch_bounds__(1024) k_packed(float *out, float a, float b)
{
    unsigned long long r[8], bb, aa;
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a), "f"(a));
#pragma unroll
    for (int i = 0; i < 8; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(r[i]) : "f"(a + i + threadIdx.x), "f"(a - i));
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(bb));
            if (OP == 1) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(bb));
            if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(bb), "l"(aa));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half the instructions packed fp32, half integer: do they share an issue slot budget only?
__global__ void __launch_bounds__(1024) k_mixed(float *out, float a, float b, int c)
{
    float r[4]; int q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { r[i] = a + i + threadIdx.x; q[i] = c + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(b));
            asm volatile("xor.b32 %0, %0, %1;" : "+r"(q[i]) : "r"(c));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += r[i] + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int WIDTH>
__global__ void __launch_bounds__(1024) k_lds(float *out)
{
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float acc = 0;
    const int base = threadIdx.x & 1023;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int idx = (base + i * 32 + it) & 1023;
            if (WIDTH == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(((float *)sm) + idx))); acc += v; }
            if (WIDTH == 8) { float v, w; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"((unsigned)__cvta_generic_to_shared(((float2 *)sm) + idx))); acc += v + w; }
            if (WIDTH == 16) { float v, w, x, y; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y) : "r"((unsigned)__cvta_generic_to_shared(sm + idx))); acc += v + w + x + y; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static void run(const char *name, F launch, double instr_per_thread, int threads, int blocks)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double warp_instr = instr_per_thread * threads / 32.0 * blocks;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at nominal %d MHz)\n", name, ms, warp_instr / cycles / sms, clk / 1000);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int T = 1024, B = sms * 2;
    float *out; cudaMalloc(&out, sizeof(float) * T * B);
    const double n = 8.0 * ITERS;
    run("FADD", [&] { k_scalar<0><<<B, T>>>(out, 1.f, 1e-3f); }, n, T, B);
    run("FMUL", [&] { k_scalar<1><<<B, T>>>(out, 1.f, 1.0001f); }, n, T, B);
    run("FFMA", [&] { k_scalar<2><<<B, T>>>(out, 1.f, 1.0001f); }, n, T, B);
    run("FADD2 (add.f32x2)", [&] { k_packed<0><<<B, T>>>(out, 1.f, 1e-3f); }, n, T, B);
    run("FMUL2 (mul.f32x2)", [&] { k_packed<1><<<B, T>>>(out, 1.f, 1.0001f); }, n, T, B);
    run("FFMA2 (fma.f32x2)", [&] { k_packed<2><<<B, T>>>(out, 1.f, 1.0001f); }, n, T, B);
    run("FADD + LOP3 interleaved", [&] { k_mixed<<<B, T>>>(out, 1.f, 1e-3f, 12345); }, n, T, B);
    run("LDS.32", [&] { k_lds<4><<<B, T>>>(out); }, n, T, B);
    run("LDS.64", [&] { k_lds<8><<<B, T>>>(out); }, n, T, B);
    run("LDS.128", [&] { k_lds<16><<<B, T>>>(out); }, n, T, B);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
