// ptx_helpers.cuh -- thin inline-PTX wrappers for sm_100a: mbarrier, 1-D bulk async copy (TMA),
// cache-hinted loads/stores, acquire/release flags.  No CUTLASS dependency.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nskptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Make barrier initialisation visible to the async (TMA) proxy.
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU (a hung box is a lost box); after ~2^26
// failed probes (seconds) the kernel traps and the host sees a launch failure.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// ---- 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP) ----
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// Same with an L2 eviction-priority policy (createpolicy result).
__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                              uint64_t *bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---- named barrier among a subset of warps ------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- global flags (device-scope release / acquire) ----------------------------------------------
__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Release-ordered increment without a return value (SASS RED): publishes every write that happens-before it.
__device__ __forceinline__ void red_release_gpu_add(int *p, int v)
{
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Plain device-scope increment; after a __threadfence() it completes a release pattern.
__device__ __forceinline__ void red_relaxed_gpu_add(int *p, int v)
{
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Coherent (L2) load of data another CTA of the same grid has just published.
__device__ __forceinline__ double ld_cg_f64(const double *p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

}  // namespace nskptx

namespace nskptx {
// Orders this thread's prior generic-proxy view of GLOBAL memory (here: data another CTA published and
// this thread acquired) before its subsequent async-proxy (bulk copy / TMA) reads of that memory.
__device__ __forceinline__ void fence_proxy_async_global()
{
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
}  // namespace nskptx

namespace nskptx {
// CTA-scope release / acquire on a shared-memory word (monotone counters between warps of one CTA)
__device__ __forceinline__ void red_release_cta_shared_add(unsigned int *p, unsigned int v)
{
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_cta_shared_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
}  // namespace nskptx
