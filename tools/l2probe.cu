// l2probe.cu -- does it matter WHICH SM reads L2-resident data?  B200 is two dies with half of the L2 each; this probe
// reads an L2-sized buffer repeatedly, either with every CTA re-reading its own slice (the pattern of tools/bwprobe.cu)
// or with the slices rotating over the CTAs from pass to pass (every pass a slice is read by a different SM), and
// a producer/consumer variant in which each slice is written by one CTA and read by another.  Loads bypass L1 (ld.cg).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/l2probe tools/l2probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint4 ld_cg(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// pass r: CTA b reads slice (b + r * shift) % grid; a slice = n16 / grid consecutive uint4
__global__ void __launch_bounds__(512) read_rot(const uint4 *buf, size_t n16, int reps, int shift, unsigned *sink)
{
    unsigned acc = 0;
    const size_t per = n16 / gridDim.x;
    for (int r = 0; r < reps; r++) {
        const size_t sl = ((size_t)blockIdx.x + (size_t)r * shift) % gridDim.x;
        const uint4 *p = buf + sl * per;
        size_t i = threadIdx.x;
        for (; i + 3 * blockDim.x < per; i += 4 * blockDim.x) {
            uint4 a = ld_cg(p + i), b = ld_cg(p + i + blockDim.x), c = ld_cg(p + i + 2 * blockDim.x), d = ld_cg(p + i + 3 * blockDim.x);
            acc ^= a.x ^ b.y ^ c.z ^ d.w;
        }
        for (; i < per; i += blockDim.x) acc ^= ld_cg(p + i).x;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// one pass: CTA b WRITES slice b, then (after a grid-wide barrier by kernel boundary) the reader kernel reads with a shift
__global__ void __launch_bounds__(512) write_own(uint4 *buf, size_t n16, unsigned v)
{
    const size_t per = n16 / gridDim.x;
    uint4 *p = buf + (size_t)blockIdx.x * per;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) p[i] = make_uint4(v, v + 1, v + 2, v + 3);
}

// latency of a dependent chain of L2 hits from one thread of every CTA; reports cycles per load per SM
__global__ void chase(const unsigned *buf, int steps, unsigned start_stride, unsigned long long *out, unsigned *smid_out)
{
    if (threadIdx.x != 0) return;
    unsigned idx = (blockIdx.x * start_stride) & 0xffff;
    unsigned long long t0 = clock64();
    for (int s = 0; s < steps; s++) {
        unsigned v;
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(buf + idx));
        idx = v;
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x] = (t1 - t0) / steps + (idx == 0xffffffffu);
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    smid_out[blockIdx.x] = sm;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s sm=%d l2=%d MB\n", prop.name, sms, prop.l2CacheSize >> 20);
    unsigned *sink;
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int mb : {16, 32, 48, 64}) {
        const size_t bytes = (size_t)mb << 20, n16 = bytes / 16;
        uint4 *buf;
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 1, bytes));
        for (int cps : {1, 2}) {
            const int grid = sms * cps;
            for (int shift : {0, 1, 37, grid / 2}) {
                const int reps = 200;
                read_rot<<<grid, 512>>>(buf, n16, 3, shift, sink);
                CK(cudaEventRecord(e0));
                read_rot<<<grid, 512>>>(buf, n16, reps, shift, sink);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("read  %3d MB ctas/sm=%d shift=%3d : %8.1f GB/s\n", mb, cps, shift, (double)bytes * reps / ms / 1e6);
            }
        }
        // producer / consumer: write a pass, read it with a shift, alternating kernels (the data is L2-resident and dirty)
        for (int shift : {0, 1, 37}) {
            const int grid = sms * 2, reps = 50;
            float tw = 0, tr = 0;
            for (int r = 0; r < reps + 2; r++) {
                CK(cudaEventRecord(e0));
                write_own<<<grid, 512>>>(buf, n16, (unsigned)r);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r >= 2) tw += ms;
                CK(cudaEventRecord(e0));
                read_rot<<<grid, 512>>>(buf, n16, 1, shift, sink);  // reps = 1: slice (b + 0) -> use shift through blockIdx offset below
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r >= 2) tr += ms;
            }
            printf("w->r  %3d MB (reader = writer CTA index; placement may differ) : write %8.1f GB/s, read %8.1f GB/s\n", mb,
                   (double)bytes * reps / tw / 1e6, (double)bytes * reps / tr / 1e6);
            break;
        }
        CK(cudaFree(buf));
    }
    // pointer chase: per-SM L2 hit latency to 16 different 2 KB-aligned cells -> near / far die pattern
    {
        const int cells = 16, words = 65536;
        unsigned *buf;
        CK(cudaMalloc(&buf, (size_t)cells * words * 4));
        unsigned long long *out;
        unsigned *smid;
        CK(cudaMalloc(&out, 8 * sms));
        CK(cudaMalloc(&smid, 4 * sms));
        unsigned long long *h = (unsigned long long *)malloc(8 * sms * cells);
        unsigned *hs = (unsigned *)malloc(4 * sms);
        unsigned *hb = (unsigned *)malloc((size_t)words * 4);
        for (int c = 0; c < cells; c++) {
            // a chain confined to ONE 2 KB cell (512 words): the cell lives on one die
            for (int i = 0; i < words; i++) hb[i] = (unsigned)((i & ~511) | ((i + 33) & 511));
            CK(cudaMemcpy(buf + (size_t)c * words, hb, (size_t)words * 4, cudaMemcpyHostToDevice));
            chase<<<sms, 32>>>(buf + (size_t)c * words, 2000, 0, out, smid);  // every CTA chases inside cell 0 of this block
            chase<<<sms, 32>>>(buf + (size_t)c * words, 4000, 0, out, smid);
            CK(cudaMemcpy(h + (size_t)c * sms, out, 8 * sms, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hs, smid, 4 * sms, cudaMemcpyDeviceToHost));
        }
        printf("L2 hit latency (cycles) per SM for %d cells of 2 KB:\n", cells);
        for (int b = 0; b < sms; b++) {
            printf("sm %3u:", hs[b]);
            for (int c = 0; c < cells; c++) printf(" %4llu", h[(size_t)c * sms + b]);
            printf("\n");
        }
    }
    printf("done\n");
    return 0;
}
