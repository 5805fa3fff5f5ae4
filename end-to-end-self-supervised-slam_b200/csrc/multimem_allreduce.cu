// All-reduce(mean) of the depth network's adaptation-gradient bucket through the NVSwitch (NVLS): the one collective of the
// data-parallel hot path (~57 MB fp32 per refinement step), written as our own kernel instead of an NCCL call.
//
// The bucket lives in symmetric memory (every rank's copy mapped into one multicast object; torch.distributed._symmetric_memory does
// the allocation and the rendezvous -- plumbing).  Rank r owns the r-th slice: one `multimem.ld_reduce` per 16 bytes makes the SWITCH
// add the eight ranks' values and return the sum, one `multimem.st` writes the mean back to all ranks at once.  Per rank that is
// numel/world loads + numel/world stores over NVLink instead of a ring's 2 (world-1)/world numel, no intermediate buffer, and a
// handful of CTAs: NCCL's ring kernel for the same bucket holds 16-32 SMs of an issue-bound neighbour kernel for ~0.3 ms (measured:
// the sweep slows from 3.18 to 3.45 ms per step at 8 GPUs), this one a few CTAs for the time the switch needs.
// Ordering: the caller brackets the launch with two cross-rank barriers on the same stream (all gradients written before; all slices
// reduced after); inside, the accesses are relaxed.
#include "common.cuh"

namespace e2e {

constexpr int MM_UNROLL = 8;       // independent switch round trips per thread (one round trip is ~2-3 us: latency, not bandwidth, bounds a thread)

__global__ void __launch_bounds__(512) multimem_allreduce_avg_kernel(float *mc, long long begin4, long long end4, float scale)
{
    const long long stride = (long long)gridDim.x * 512;
    for (long long i0 = begin4 + (long long)blockIdx.x * 512 + threadIdx.x; i0 < end4; i0 += stride * MM_UNROLL) {
        float4 v[MM_UNROLL];
#pragma unroll
        for (int u = 0; u < MM_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < end4)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                             : "l"(mc + 4 * i)
                             : "memory");
        }
#pragma unroll
        for (int u = 0; u < MM_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < end4)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(v[u].x * scale), "f"(v[u].y * scale),
                             "f"(v[u].z * scale), "f"(v[u].w * scale)
                             : "memory");
        }
    }
}

}  // namespace e2e

using namespace e2e;

extern "C" int e2e_multimem_allreduce_avg(void *multicast_ptr, long long numel, int rank, int world, int ctas, void *stream)
{
    E2E_REQUIRE(multicast_ptr && numel > 0 && world >= 1 && rank >= 0 && rank < world, "multimem_allreduce: bad arguments");
    E2E_REQUIRE(numel % 4 == 0 && (((uintptr_t)multicast_ptr) & 15u) == 0, "multimem_allreduce: the bucket must be a multiple of 4 floats and 16-byte aligned");
    const long long n4 = numel / 4, per = (n4 + world - 1) / world;
    const long long begin4 = per * rank, end4 = (begin4 + per < n4) ? begin4 + per : n4;
    if (begin4 >= end4) return 0;
    if (ctas < 1) ctas = 8;              // measured: 8, 16 and 32 CTAs take the same time (the switch, not the SMs, is the limit)
    long long need = (end4 - begin4 + 511) / 512;
    if (need < ctas) ctas = (int)need;
    multimem_allreduce_avg_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>((float *)multicast_ptr, begin4, end4, 1.0f / (float)world);
    count_launch();
    return finish_launch("multimem_allreduce_avg_kernel");
}
