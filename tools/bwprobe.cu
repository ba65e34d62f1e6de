// bwprobe.cu -- measures what the roofline arguments in DESIGN.md rest on: HBM copy / read bandwidth
// and L2-resident read bandwidth (plain LDG.128 and TMA bulk copies) on the GPU it runs on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/bwprobe tools/bwprobe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../navierstokes_b200/csrc/ptx_helpers.cuh"
using namespace nskptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(512) read_kernel(const uint4 *__restrict__ buf, size_t n16, int reps, unsigned *sink)
{
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n16; i += 4 * stride) {
            uint4 a = __ldg(buf + i), b = __ldg(buf + i + stride), c = __ldg(buf + i + 2 * stride), d = __ldg(buf + i + 3 * stride);
            acc ^= a.x ^ b.y ^ c.z ^ d.w;
        }
        for (; i < n16; i += stride) acc ^= __ldg(buf + i).x;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(512) copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride), d = __ldg(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = __ldg(src + i);
}

// TMA bulk reads into a shared-memory ring: chunk bytes per copy, STAGES in flight per CTA.
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(64) tma_read_kernel(const unsigned char *buf, size_t bytes, int reps, unsigned *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t)CHUNK * STAGES);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = bytes / CHUNK;
    unsigned acc = 0;
    uint32_t phase_bits = 0;  // per stage: parity of the next completion to wait for (persists across reps)
    for (int r = 0; r < reps; r++) {
        // issue-ahead ring: wait for chunk it-STAGES before reusing its slot
        uint32_t it = 0;
        size_t c = blockIdx.x;
        for (; c < nchunks; c += gridDim.x, ++it) {
            int s = it % STAGES;
            if (it >= STAGES) {
                mbar_wait(&full[s], (phase_bits >> s) & 1);
                phase_bits ^= (1u << s);
                acc ^= *reinterpret_cast<volatile unsigned *>(sm + (size_t)s * CHUNK);
            }
            mbar_arrive_expect_tx(&full[s], CHUNK);
            bulk_g2s(sm + (size_t)s * CHUNK, buf + c * CHUNK, CHUNK, &full[s]);
        }
        // drain
        uint32_t total = it;
        for (uint32_t j = (total > STAGES ? total - STAGES : 0); j < total; j++) {
            int s = j % STAGES;
            mbar_wait(&full[s], (phase_bits >> s) & 1);
            phase_bits ^= (1u << s);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm=%d l2=%d MB smem_optin=%zu\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, p.sharedMemPerBlockOptin);
    const size_t GiB = 1ull << 30;
    unsigned char *a, *b; unsigned *sink;
    CK(cudaMalloc(&a, 2 * GiB)); CK(cudaMalloc(&b, 2 * GiB)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, 2 * GiB)); CK(cudaMemset(b, 2, 2 * GiB));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int sms = p.multiProcessorCount;

    for (int cps : {2, 4}) {
        float best = 1e9;
        for (int it = 0; it < 6; it++) {
            CK(cudaEventRecord(e0));
            copy_kernel<<<sms * cps, 512>>>((const uint4 *)a, (uint4 *)b, 2 * GiB / 16);
            CK(cudaEventRecord(e1));
            float ms = time_ms(e0, e1); if (it && ms < best) best = ms;
        }
        printf("copy   2GiB->2GiB  ctas/sm=%d : %.3f ms  %.1f GB/s (read+write)\n", cps, best, 4.0 * GiB / best / 1e6);
    }
    {
        float best = 1e9;
        for (int it = 0; it < 6; it++) {
            CK(cudaEventRecord(e0));
            CK(cudaMemcpyAsync(b, a, 2 * GiB, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(e1));
            float ms = time_ms(e0, e1); if (it && ms < best) best = ms;
        }
        printf("cudaMemcpy D2D 2GiB         : %.3f ms  %.1f GB/s (read+write)\n", best, 4.0 * GiB / best / 1e6);
    }
    size_t sizes_mb[] = {8, 16, 32, 48, 64, 80, 96, 112, 128, 192, 256, 1024, 2048};
    for (size_t mb : sizes_mb) {
        size_t bytes = mb << 20;
        int reps = (int)((8 * GiB) / bytes); if (reps < 2) reps = 2; if (reps > 400) reps = 400;
        for (int cps : {2, 4}) {
            float best = 1e9;
            for (int it = 0; it < 4; it++) {
                CK(cudaEventRecord(e0));
                read_kernel<<<sms * cps, 512>>>((const uint4 *)a, bytes / 16, reps, sink);
                CK(cudaEventRecord(e1));
                float ms = time_ms(e0, e1); if (it && ms < best) best = ms;
            }
            printf("read LDG.128  %5zu MB x%3d ctas/sm=%d : %8.3f ms  %8.1f GB/s\n", mb, reps, cps, best, (double)bytes * reps / best / 1e6);
        }
    }
    {
        constexpr int CHUNK = 16384, STAGES = 8;
        int smem = CHUNK * STAGES + 8 * STAGES + 64;
        CK(cudaFuncSetAttribute(tma_read_kernel<CHUNK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (size_t mb : {32, 64, 96, 2048}) {
            size_t bytes = (size_t)mb << 20;
            int reps = (int)((8 * GiB) / bytes); if (reps < 2) reps = 2; if (reps > 400) reps = 400;
            float best = 1e9;
            for (int it = 0; it < 4; it++) {
                CK(cudaEventRecord(e0));
                tma_read_kernel<CHUNK, STAGES><<<sms, 64, smem>>>(a, bytes, reps, sink);
                CK(cudaEventRecord(e1));
                float ms = time_ms(e0, e1); if (it && ms < best) best = ms;
            }
            printf("read TMA bulk %5zu MB x%3d 1 cta/sm %dx%dKB : %8.3f ms  %8.1f GB/s\n", (size_t)mb, reps, STAGES, CHUNK >> 10, best, (double)bytes * reps / best / 1e6);
        }
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
