// context.cu -- context, memory and error plumbing of the nsk C ABI (include/nsk.h).
#include <stdarg.h>
#include <string.h>

#include "nsk_internal.h"

static thread_local std::string g_last_error;  // failures that happen without a context

void nsk_set_error(nsk_ctx_t ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    g_last_error = buf;
}

NSK_API int nsk_version(void) { return NSK_VERSION; }

NSK_API const char *nsk_strerror(int s)
{
    switch (s) {
        case NSK_OK: return "ok";
        case NSK_ERR_INVALID: return "invalid argument";
        case NSK_ERR_CUDA: return "CUDA error";
        case NSK_ERR_NO_DEVICE: return "no sm_100 GPU available (there is no CPU fallback)";
        case NSK_ERR_ALLOC: return "allocation failed";
        case NSK_ERR_COMM: return "communicator error";
        case NSK_ERR_UNSUPPORTED: return "unsupported";
        case NSK_ERR_NOT_CONVERGED: return "not converged";
    }
    return "unknown status";
}

NSK_API const char *nsk_last_error(nsk_ctx_t ctx)
{
    return ctx ? ctx->last_error.c_str() : g_last_error.c_str();
}

NSK_API int nsk_ctx_create(int device, nsk_ctx_t *out)
{
    if (!out) return NSK_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        nsk_set_error(nullptr, "no CUDA device visible (%s); this library has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return NSK_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        nsk_set_error(nullptr, "device %d out of range (%d visible)", device, count);
        return NSK_ERR_INVALID;
    }
    nsk_ctx_s *c = new nsk_ctx_s();
    c->device = device;
    NSK_CUDA(nullptr, cudaSetDevice(device));
    NSK_CUDA(nullptr, cudaGetDeviceProperties(&c->prop, device));
    if (c->prop.major != 10) {
        nsk_set_error(nullptr, "device %d is sm_%d%d; the kernels are built for sm_100a only",
                      device, c->prop.major, c->prop.minor);
        delete c;
        return NSK_ERR_NO_DEVICE;
    }
    NSK_CUDA(nullptr, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    NSK_CUDA(nullptr, cudaMalloc(&c->d_partials, sizeof(double) * NSK_MAX_PARTIALS * NSK_RED_SLOTS));
    NSK_CUDA(nullptr, cudaMalloc(&c->d_ticket, sizeof(unsigned int) * 16));
    NSK_CUDA(nullptr, cudaMemset(c->d_ticket, 0, sizeof(unsigned int) * 16));
    NSK_CUDA(nullptr, cudaMalloc(&c->d_scalars, sizeof(double) * NSK_NSCALARS));
    NSK_CUDA(nullptr, cudaMemset(c->d_scalars, 0, sizeof(double) * NSK_NSCALARS));
    NSK_CUDA(nullptr, cudaMallocHost(&c->h_scalars, sizeof(double) * NSK_NSCALARS));
    *out = c;
    return NSK_OK;
}

NSK_API int nsk_ctx_destroy(nsk_ctx_t c)
{
    if (!c) return NSK_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    nsk_comm_destroy(c);
    nsk_ipc_close_all(c);
    for (int i = 0; i < 8; i++) if (c->d_stage[i]) cudaFree(c->d_stage[i]);
    if (c->d_flush) cudaFree(c->d_flush);
    cudaFree(c->d_partials);
    cudaFree(c->d_ticket);
    cudaFree(c->d_scalars);
    cudaFreeHost(c->h_scalars);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < NSK_MAX_K; i++)
        if (c->copy_event[i]) cudaEventDestroy(c->copy_event[i]);
    delete c;
    return NSK_OK;
}

NSK_API int nsk_ctx_set_stream(nsk_ctx_t c, void *s)
{
    if (!c) return NSK_ERR_INVALID;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return NSK_OK;
}

NSK_API void *nsk_ctx_get_stream(nsk_ctx_t c) { return c ? (void *)c->stream : nullptr; }

NSK_API int nsk_ctx_sync(nsk_ctx_t c)
{
    if (!c) return NSK_ERR_INVALID;
    NSK_CUDA(c, cudaStreamSynchronize(c->stream));
    return NSK_OK;
}

NSK_API uint64_t nsk_ctx_launch_count(nsk_ctx_t c) { return c ? c->launches : 0; }

NSK_API int nsk_ctx_device_info(nsk_ctx_t c, int *sm_count, int64_t *l2_bytes, int *smem_optin,
                                int64_t *hbm_bytes)
{
    if (!c) return NSK_ERR_INVALID;
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = c->prop.l2CacheSize;
    if (smem_optin) *smem_optin = (int)c->prop.sharedMemPerBlockOptin;
    if (hbm_bytes) *hbm_bytes = (int64_t)c->prop.totalGlobalMem;
    return NSK_OK;
}

NSK_API int nsk_ctx_set_option(nsk_ctx_t c, const char *name, int64_t v)
{
    if (!c || !name) return NSK_ERR_INVALID;
    if (!strcmp(name, "spmv_kernel")) c->opt.spmv_kernel = v;
    else if (!strcmp(name, "spmv_ctas_per_sm")) c->opt.spmv_ctas_per_sm = v;
    else if (!strcmp(name, "mpk_kernel")) c->opt.mpk_kernel = v;
    else if (!strcmp(name, "stream_variant")) c->opt.stream_variant = v;
    else if (!strcmp(name, "wave_slack_pct")) c->opt.wave_slack_pct = v;
    else if (!strcmp(name, "wave_l2_pct")) c->opt.wave_l2_pct = v;
    else if (!strcmp(name, "packed_variant")) c->opt.packed_variant = v;
    else if (!strcmp(name, "pipe_bp_global")) c->opt.pipe_bp_global = v;
    else if (!strcmp(name, "pipe_w0_pct")) c->opt.pipe_w0_pct = v;
    else if (!strcmp(name, "pk_timing")) c->opt.pk_timing = v;
    else if (!strcmp(name, "pk_flags")) c->opt.pk_flags = v;
    else if (!strcmp(name, "host_overlap")) c->opt.host_overlap = v;
    else if (!strcmp(name, "stream_exact_kind")) c->opt.stream_exact_kind = v;
    else if (!strcmp(name, "pipe_interleave")) c->opt.pipe_interleave = v;
    else if (!strcmp(name, "halo_push")) c->opt.halo_push = v;
    else if (!strcmp(name, "local_reductions")) c->opt.local_reductions = v;
    else if (!strcmp(name, "gram_wide")) c->opt.gram_wide = v;
    else if (!strcmp(name, "mpk_auto_explicit")) c->opt.mpk_auto_explicit = v;
    else if (!strcmp(name, "scg_update_wide")) c->opt.scg_update_wide = v;
    else if (!strcmp(name, "bcsr_batch")) c->opt.bcsr_batch = v;
    else if (!strcmp(name, "sell_chunk")) c->opt.sell_chunk = v;
    else if (!strcmp(name, "sell_geom")) c->opt.sell_geom = v;
    else if (!strcmp(name, "sell_ctas_per_sm")) c->opt.sell_ctas_per_sm = v;
    else if (!strcmp(name, "sell_max_ctas")) c->opt.sell_max_ctas = v;
    else if (!strcmp(name, "sell_flags")) c->opt.sell_flags = v;
    else if (!strcmp(name, "sell_pf_dist")) c->opt.sell_pf_dist = v;
    else if (!strcmp(name, "sell_stream")) c->opt.sell_stream = v;
    else if (!strcmp(name, "sell_rows")) c->opt.sell_rows = v;
    else if (!strcmp(name, "sell_tma")) c->opt.sell_tma = v;
    else {
        nsk_set_error(c, "unknown option '%s'", name);
        return NSK_ERR_INVALID;
    }
    return NSK_OK;
}

NSK_API int nsk_ctx_query(nsk_ctx_t c, const char *name, int64_t *value)
{
    if (!c || !name || !value) return NSK_ERR_INVALID;
    if (!strcmp(name, "last_spmv_kernel")) *value = c->last_spmv;
    else if (!strcmp(name, "last_mpk_strategy")) *value = c->last_mpk;
    else if (!strcmp(name, "launches")) *value = (int64_t)c->launches;
    else if (!strcmp(name, "sell_uniform_width")) *value = c->last_sell[0];
    else if (!strcmp(name, "sell_reach")) *value = c->last_sell[1];
    else if (!strcmp(name, "sell_lead")) *value = c->last_sell[2];
    else if (!strcmp(name, "sell_grid")) *value = c->last_sell[3];
    else if (!strcmp(name, "sell_identity_tiles")) *value = c->last_sell[4];
    else if (!strcmp(name, "sell_ntiles")) *value = c->last_sell[5];
    else if (!strcmp(name, "sell_staged")) *value = c->last_sell[6];
    else if (!strcmp(name, "sell_ngroups")) *value = c->last_sell[7];
    else {
        nsk_set_error(c, "unknown query '%s'", name);
        return NSK_ERR_INVALID;
    }
    return NSK_OK;
}

// ---- events ---------------------------------------------------------------------------------
NSK_API int nsk_event_create(nsk_ctx_t c, void **ev)
{
    if (!c || !ev) return NSK_ERR_INVALID;
    cudaEvent_t e;
    NSK_CUDA(c, cudaEventCreate(&e));
    *ev = (void *)e;
    return NSK_OK;
}

NSK_API int nsk_event_destroy(nsk_ctx_t c, void *ev)
{
    if (!c) return NSK_ERR_INVALID;
    if (ev) NSK_CUDA(c, cudaEventDestroy((cudaEvent_t)ev));
    return NSK_OK;
}

NSK_API int nsk_event_record(nsk_ctx_t c, void *ev)
{
    if (!c || !ev) return NSK_ERR_INVALID;
    NSK_CUDA(c, cudaEventRecord((cudaEvent_t)ev, c->stream));
    return NSK_OK;
}

NSK_API int nsk_event_elapsed_ms(nsk_ctx_t c, void *start, void *stop, float *ms)
{
    if (!c || !start || !stop || !ms) return NSK_ERR_INVALID;
    NSK_CUDA(c, cudaEventSynchronize((cudaEvent_t)stop));
    NSK_CUDA(c, cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return NSK_OK;
}

// ---- memory ---------------------------------------------------------------------------------
NSK_API int nsk_malloc(nsk_ctx_t c, size_t bytes, void **p)
{
    if (!c || !p) return NSK_ERR_INVALID;
    *p = nullptr;
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        nsk_set_error(c, "cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
        return NSK_ERR_ALLOC;
    }
    return NSK_OK;
}

NSK_API int nsk_free(nsk_ctx_t c, void *p)
{
    if (!c) return NSK_ERR_INVALID;
    if (p) NSK_CUDA(c, cudaFree(p));
    return NSK_OK;
}

NSK_API int nsk_host_alloc(nsk_ctx_t c, size_t bytes, void **p)
{
    if (!c || !p) return NSK_ERR_INVALID;
    cudaError_t e = cudaMallocHost(p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        nsk_set_error(c, "cudaMallocHost(%zu) -> %s", bytes, cudaGetErrorString(e));
        return NSK_ERR_ALLOC;
    }
    return NSK_OK;
}

NSK_API int nsk_host_free(nsk_ctx_t c, void *p)
{
    if (!c) return NSK_ERR_INVALID;
    if (p) NSK_CUDA(c, cudaFreeHost(p));
    return NSK_OK;
}

NSK_API int nsk_memcpy(nsk_ctx_t c, void *dst, const void *src, size_t bytes, int kind)
{
    if (!c || (bytes && (!dst || !src))) return NSK_ERR_INVALID;
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice
                     : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (bytes) NSK_CUDA(c, cudaMemcpyAsync(dst, src, bytes, k, c->stream));
    return NSK_OK;
}

NSK_API int nsk_memset0(nsk_ctx_t c, void *p, size_t bytes)
{
    if (!c) return NSK_ERR_INVALID;
    if (bytes) NSK_CUDA(c, cudaMemsetAsync(p, 0, bytes, c->stream));
    return NSK_OK;
}

int nsk_stage(nsk_ctx_t c, int slot, size_t bytes, void **p)
{
    if (slot < 0 || slot >= 8) return NSK_ERR_INVALID;
    if (c->stage_bytes[slot] < bytes) {
        if (c->d_stage[slot]) {
            NSK_CUDA(c, cudaStreamSynchronize(c->stream));
            NSK_CUDA(c, cudaFree(c->d_stage[slot]));
            c->d_stage[slot] = nullptr;
            c->stage_bytes[slot] = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&c->d_stage[slot], want);
        if (e != cudaSuccess) {
            nsk_set_error(c, "staging cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
            return NSK_ERR_ALLOC;
        }
        c->stage_bytes[slot] = want;
    }
    *p = c->d_stage[slot];
    return NSK_OK;
}

__global__ void nsk_flush_kernel(float4 *buf, size_t n4, float v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) buf[i] = make_float4(v, v, v, v);
}

NSK_API int nsk_flush_l2(nsk_ctx_t c)
{
    if (!c) return NSK_ERR_INVALID;
    if (!c->d_flush) {
        c->flush_bytes = (size_t)c->prop.l2CacheSize * 2 + (64u << 20);
        cudaError_t e = cudaMalloc(&c->d_flush, c->flush_bytes);
        if (e != cudaSuccess) {
            nsk_set_error(c, "flush cudaMalloc -> %s", cudaGetErrorString(e));
            return NSK_ERR_ALLOC;
        }
    }
    static float v = 0.f;
    v += 1.f;
    nsk_flush_kernel<<<c->prop.multiProcessorCount * 8, 256, 0, c->stream>>>(
        (float4 *)c->d_flush, c->flush_bytes / 16, v);
    c->launches++;
    NSK_CUDA(c, cudaGetLastError());
    return NSK_OK;
}
