// comm.cpp -- NCCL communicator over NVLink for the multi-GPU path (one process per GPU).
//
// NCCL is loaded with dlopen at first use: when the process already holds torch's bundled
// libnccl.so.2 (torch.distributed) the same image is reused, otherwise the system library is
// opened.  Nothing here is needed (or touched) for single-GPU use.
#include <dlfcn.h>
#include <string.h>

#include "nsk_internal.h"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[NSK_UNIQUE_ID_BYTES]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };  // ncclDataType_t: ncclDouble
enum { ncclSum = 0 };      // ncclRedOp_t

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl(nsk_ctx_t ctx)
{
    if (g_nccl.handle) return NSK_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        nsk_set_error(ctx, "cannot dlopen libnccl.so.2: %s", dlerror());
        return NSK_ERR_COMM;
    }
#define LOAD(field, sym)                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                             \
    if (!g_nccl.field) {                                                   \
        nsk_set_error(ctx, "libnccl lacks %s", sym);                       \
        return NSK_ERR_COMM;                                               \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_nccl.handle = h;
    return NSK_OK;
}

struct nsk_comm_s {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

#define NSK_NCCL(ctx, call)                                                              \
    do {                                                                                 \
        ncclResult_t r__ = (call);                                                       \
        if (r__ != 0) {                                                                  \
            nsk_set_error((ctx), "%s:%d %s -> %s", __FILE__, __LINE__, #call,            \
                          g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl");  \
            return NSK_ERR_COMM;                                                         \
        }                                                                                \
    } while (0)

NSK_API int nsk_comm_unique_id(void *id128)
{
    if (!id128) return NSK_ERR_INVALID;
    NSK_TRY(load_nccl(nullptr));
    ncclUniqueId id;
    NSK_NCCL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, NSK_UNIQUE_ID_BYTES);
    return NSK_OK;
}

NSK_API int nsk_comm_init(nsk_ctx_t ctx, int nranks, int rank, const void *id128)
{
    if (!ctx || !id128) return NSK_ERR_INVALID;
    NSK_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    NSK_TRY(load_nccl(ctx));
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->comm) nsk_comm_destroy(ctx);
    nsk_comm_s *c = new nsk_comm_s();
    c->nranks = nranks;
    c->rank = rank;
    ncclUniqueId id;
    memcpy(&id, id128, NSK_UNIQUE_ID_BYTES);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) {
        nsk_set_error(ctx, "ncclCommInitRank -> %s", g_nccl.GetErrorString(r));
        delete c;
        return NSK_ERR_COMM;
    }
    ctx->comm = c;
    return NSK_OK;
}

NSK_API int nsk_comm_destroy(nsk_ctx_t ctx)
{
    if (!ctx || !ctx->comm) return NSK_OK;
    if (ctx->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
    return NSK_OK;
}

bool nsk_comm_active(nsk_ctx_t ctx) { return ctx && ctx->comm && ctx->comm->nranks > 1; }

int nsk_comm_rank(nsk_ctx_t ctx) { return ctx && ctx->comm ? ctx->comm->rank : 0; }
int nsk_comm_size(nsk_ctx_t ctx) { return ctx && ctx->comm ? ctx->comm->nranks : 1; }

NSK_API int nsk_comm_allreduce_sum(nsk_ctx_t ctx, double *dbuf, int count)
{
    if (!ctx) return NSK_ERR_INVALID;
    if (!nsk_comm_active(ctx) || count <= 0) return NSK_OK;
    NSK_NCCL(ctx, g_nccl.AllReduce(dbuf, dbuf, (size_t)count, ncclFloat64, ncclSum, ctx->comm->comm, ctx->stream));
    return NSK_OK;
}

int nsk_comm_allreduce_slots(nsk_ctx_t ctx, int slot0, int count)
{
    if (!nsk_comm_active(ctx)) return NSK_OK;
    return nsk_comm_allreduce_sum(ctx, ctx->d_scalars + slot0, count);
}

int nsk_comm_sendrecv(nsk_ctx_t ctx, int npeers, const int *peer, const double *const *sendbuf,
                      const int *sendcount, double *const *recvbuf, const int *recvcount)
{
    if (!nsk_comm_active(ctx)) return NSK_OK;
    NSK_NCCL(ctx, g_nccl.GroupStart());
    for (int i = 0; i < npeers; i++) {
        if (sendcount[i] > 0)
            NSK_NCCL(ctx, g_nccl.Send(sendbuf[i], (size_t)sendcount[i], ncclFloat64, peer[i], ctx->comm->comm, ctx->stream));
        if (recvcount[i] > 0)
            NSK_NCCL(ctx, g_nccl.Recv(recvbuf[i], (size_t)recvcount[i], ncclFloat64, peer[i], ctx->comm->comm, ctx->stream));
    }
    NSK_NCCL(ctx, g_nccl.GroupEnd());
    return NSK_OK;
}
