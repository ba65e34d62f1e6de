// dist.cu -- device side of the row-partitioned operator: halo exchange (pack kernel + grouped
// NCCL point-to-point over NVLink) and matrix powers with redundant ghost levels.
//
// The reference is single-process; this is the multi-GPU layer BASELINE.json's north
// star asks for: one process per GPU, the depth-k halo exchanged ONCE per matrix-powers call, ghost
// levels recomputed redundantly on shrinking row prefixes (plan: dist_plan.cpp).
#include <algorithm>
#include <cstring>

#include "dist_plan.h"
#include "nsk_internal.h"

int nsk_mpk_local(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode,
                  const int *level_rows);  // mpk.cu
int nsk_mpk_local2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                   double *const *d_levels2, nsk_mode mode, const int *level_rows);  // mpk.cu
void nsk_wave_set_block_extents(nsk_csr_t A, const int *ptrow, const int *indcol);
int nsk_comm_rank(nsk_ctx_t ctx);
int nsk_comm_size(nsk_ctx_t ctx);

// ---- halo PUSH over NVLink peer memory --------------------------------------------------------------------------
// Vectors created through nsk_dist_vector_register are mapped into the neighbours' address spaces (CUDA IPC, the handles
// travel through the host layer once).  For such a vector the depth-k halo of a matrix-powers call is ONE kernel: every
// rank gathers the entries its neighbours need and stores them straight into the neighbours' copies of the vector --
// into the ghost slots themselves, so there is no send buffer, no receive buffer, no unpack and no NCCL call on the
// path -- then raises an arrival flag in each neighbour's memory (fence.sys + st.release.sys).  A one-warp kernel on
// the receiving side waits for its neighbours' flags (ld.acquire.sys); after the powers kernel has consumed the ghost
// entries an acknowledgement flag goes back, which the neighbour's NEXT push into the same vector waits for.  Flags are
// epochs (one counter per registered vector), nothing is ever reset.
constexpr int PUSH_SLOTS = 64;     // registered vectors per operator
constexpr int PUSH_MAXPEERS = 8;
struct PushFlags {
    unsigned int arrive[PUSH_SLOTS][PUSH_MAXPEERS];  // [vector][peer index here]: epoch of the last halo that peer pushed
    unsigned int ack[PUSH_SLOTS][PUSH_MAXPEERS];     // [vector][peer index here]: epoch of the last halo that peer consumed
};
struct PushVector {
    double *local = nullptr;
    double *peer[PUSH_MAXPEERS] = {};  // the neighbours' copies, mapped here
    // Epochs count per neighbour and direction: a shallow exchange (depth 1) may involve fewer neighbours than a deep one,
    // and the two ends of a link agree on which exchanges used it (the sender's ring counts ARE the receiver's).
    unsigned int sent[PUSH_MAXPEERS] = {};
    unsigned int received[PUSH_MAXPEERS] = {};
    bool pending = false;              // pushed and not yet released
};

struct DistPeer {
    int rank = 0;
    int *d_send_idx = nullptr;          // local indices to pack, ring-major
    double *d_sendbuf = nullptr;
    double *d_recvbuf = nullptr;        // one contiguous message per peer lands here, then unpack_kernel scatters it
    std::vector<int> send_ring_count;   // depth
    std::vector<int> recv_ring_count;   // depth
    std::vector<int> recv_ring_start;   // depth, offset into the local vector
    // push path
    int index_at_peer = -1;             // position of this rank in the peer's own peer list
    int *d_push_dst = nullptr;          // per entry of the send list: where it lands in the PEER's local vector
    PushFlags *peer_flags = nullptr;    // the peer's flags, mapped here
};

struct nsk_dist_s {
    int n_owned = 0, depth = 0;
    std::vector<int> ring_start;  // depth + 2
    std::vector<DistPeer> peers;
    // push path
    PushFlags *d_flags = nullptr;       // this rank's flags (IPC-exported to the neighbours)
    unsigned int *d_ticket = nullptr;
    int *h_error = nullptr, *d_error = nullptr;  // host-mapped: a bounded wait expired
    std::vector<PushVector> vecs;
    bool push_ready = false;
};

// All peers in ONE launch each way (a per-ring message costs ~9 us of NCCL latency, a launch ~3 us: measured 46 us
// per depth-4 exchange with one neighbour when every ring travelled alone, tools/dist_probe.py).
constexpr int HALO_MAX_SEG = 64;  // peers x rings
struct HaloSegs {
    int nseg;
    int total;
    int begin[HALO_MAX_SEG + 1];    // prefix sums of segment lengths
    const double *src[HALO_MAX_SEG];
    double *dst[HALO_MAX_SEG];
    const int *idx[HALO_MAX_SEG];   // pack: gather indices (src = local vector); unpack: nullptr (contiguous copy)
};

__global__ void __launch_bounds__(256) halo_move_kernel(const HaloSegs S)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.total) return;
    int s = 0;
    while (s + 1 < S.nseg && i >= S.begin[s + 1]) s++;
    const int j = i - S.begin[s];
    S.dst[s][j] = S.idx[s] ? S.src[s][S.idx[s][j]] : S.src[s][j];
}

void nsk_dist_free(nsk_csr_t A)
{
    if (!A || !A->dist) return;
    for (DistPeer &p : A->dist->peers) {
        if (p.d_send_idx) cudaFree(p.d_send_idx);
        if (p.d_sendbuf) cudaFree(p.d_sendbuf);
        if (p.d_recvbuf) cudaFree(p.d_recvbuf);
        if (p.d_push_dst) cudaFree(p.d_push_dst);
    }
    if (A->dist->d_flags) cudaFree(A->dist->d_flags);
    if (A->dist->d_ticket) cudaFree(A->dist->d_ticket);
    if (A->dist->h_error) cudaFreeHost(A->dist->h_error);
    delete A->dist;
    A->dist = nullptr;
}

NSK_API int nsk_csr_owned_rows(nsk_csr_t A)
{
    if (!A) return 0;
    return A->dist ? A->dist->n_owned : A->n;
}

NSK_API int nsk_csr_create_dist(nsk_ctx_t ctx, nsk_plan_t plan, nsk_csr_t *out)
{
    if (!ctx || !plan || !out) return NSK_ERR_INVALID;
    NSK_REQUIRE(ctx, plan->finalized, "plan is not finalized");
    const int n_rows = plan->ring_start[plan->depth];
    const int n_cols = plan->ring_start[plan->depth + 1];
    nsk_csr_t A = nullptr;
    NSK_TRY(nsk_csr_create(ctx, n_rows, n_cols, (int64_t)plan->local_cols.size(), plan->ptr.data(),
                           plan->local_cols.data(), plan->vals.data(), &A));
    // Tiles must not straddle a jump in global row order, and the fused matrix-powers kernel wants to know
    // where each stored row sits in that order (ghost rings are stored AFTER the owned rows but belong
    // below / above them): record both, then redo the column extents in rank space.
    {
        const int n_owned = plan->own_end - plan->own_begin;
        std::vector<int> gid(n_rows);
        for (int i = 0; i < n_rows; i++) gid[i] = i < n_owned ? plan->own_begin + i : plan->ghost_gids[i - n_owned];
        std::vector<int> order(n_rows);
        for (int i = 0; i < n_rows; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return gid[a] < gid[b]; });
        A->row_rank.assign(n_rows, 0);
        for (int r = 0; r < n_rows; r++) A->row_rank[order[r]] = r;
        A->breaks.clear();
        for (int i = 1; i < n_rows; i++)
            if (gid[i] != gid[i - 1] + 1) A->breaks.push_back(i);
        nsk_wave_set_block_extents(A, plan->ptr.data(), plan->local_cols.data());
    }
    nsk_dist_s *D = new nsk_dist_s();
    D->n_owned = plan->own_end - plan->own_begin;
    D->depth = plan->depth;
    D->ring_start = plan->ring_start;
    A->dist = D;
    // peers = every rank we send to or receive from, ascending
    std::vector<int> ranks;
    for (auto &kv : plan->req) ranks.push_back(kv.first);
    for (auto &kv : plan->sends) ranks.push_back(kv.first);
    std::sort(ranks.begin(), ranks.end());
    ranks.erase(std::unique(ranks.begin(), ranks.end()), ranks.end());
    for (int r : ranks) {
        DistPeer P;
        P.rank = r;
        P.send_ring_count.assign(plan->depth, 0);
        P.recv_ring_count.assign(plan->depth, 0);
        P.recv_ring_start.assign(plan->depth, 0);
        auto rq = plan->req.find(r);
        if (rq != plan->req.end()) {
            P.recv_ring_count = rq->second.ring_count;
            P.recv_ring_start = rq->second.ring_local_start;
        }
        auto sd = plan->sends.find(r);
        if (sd != plan->sends.end() && !sd->second.local_idx.empty()) {
            P.send_ring_count = sd->second.ring_count;
            const size_t cnt = sd->second.local_idx.size();
            if (cudaMalloc(&P.d_send_idx, sizeof(int) * cnt) != cudaSuccess ||
                cudaMalloc(&P.d_sendbuf, sizeof(double) * cnt) != cudaSuccess) {
                nsk_set_error(ctx, "halo buffer allocation failed");
                nsk_csr_destroy(A);
                return NSK_ERR_ALLOC;
            }
            NSK_CUDA(ctx, cudaMemcpy(P.d_send_idx, sd->second.local_idx.data(), sizeof(int) * cnt, cudaMemcpyHostToDevice));
        }
        {
            size_t rc = 0;
            for (int v : P.recv_ring_count) rc += (size_t)v;
            if (rc > 0 && cudaMalloc(&P.d_recvbuf, sizeof(double) * rc) != cudaSuccess) {
                nsk_set_error(ctx, "halo buffer allocation failed");
                nsk_csr_destroy(A);
                return NSK_ERR_ALLOC;
            }
        }
        D->peers.push_back(P);
    }
    *out = A;
    return NSK_OK;
}

// ---- push kernels ---------------------------------------------------------------------------------------------------
constexpr unsigned long long PUSH_TIMEOUT_NS = 4000000000ull;
__device__ __forceinline__ unsigned long long push_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PushArgs {
    int npeers, total;
    unsigned int send_epoch[PUSH_MAXPEERS], wait_epoch[PUSH_MAXPEERS];
    const double *x;
    int begin[PUSH_MAXPEERS + 1];
    const int *send_idx[PUSH_MAXPEERS];
    const int *dst_idx[PUSH_MAXPEERS];
    double *peer_vec[PUSH_MAXPEERS];
    const unsigned int *my_ack[PUSH_MAXPEERS];   // local: the peer consumed what I pushed last time
    unsigned int *peer_arrive[PUSH_MAXPEERS];    // remote: my halo has landed
    int nwait;
    const unsigned int *my_arrive[PUSH_MAXPEERS];  // local: the neighbours' halos have landed here
    unsigned int *ticket;
    int *error;
};

__global__ void __launch_bounds__(256) halo_push_kernel(const PushArgs a)
{
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        // write-after-read: the neighbour must have consumed the previous halo it got in this vector
        int ok = 1;
        for (int j = 0; j < a.npeers; j++) {
            unsigned long long t0 = 0;
            while (ld_acquire_sys(a.my_ack[j]) + 1u < a.send_epoch[j]) {
                __nanosleep(100);
                const unsigned long long t = push_now();
                if (t0 == 0) t0 = t;
                else if (t - t0 > PUSH_TIMEOUT_NS) { ok = 0; break; }
            }
        }
        if (!ok) *a.error = 1;
        s_ok = ok;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.total) {
        int j = 0;
        while (j + 1 < a.npeers && i >= a.begin[j + 1]) j++;
        const int q = i - a.begin[j];
        a.peer_vec[j][a.dst_idx[j][q]] = a.x[a.send_idx[j][q]];  // NVLink store into the neighbour's ghost slot
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) {  // every block's stores are fenced: tell the neighbours
            __threadfence_system();
            for (int j = 0; j < a.npeers; j++) st_release_sys(a.peer_arrive[j], a.send_epoch[j]);
            *a.ticket = 0u;
            // ... and wait for theirs: the kernel (hence everything behind it on the stream) ends when the ghost slots
            // of this vector are complete.  Signalling comes first on every rank, so the waits cannot form a cycle.
            for (int j = 0; j < a.nwait; j++) {
                unsigned long long t0 = 0;
                while (ld_acquire_sys(a.my_arrive[j]) < a.wait_epoch[j]) {
                    __nanosleep(100);
                    const unsigned long long t = push_now();
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > PUSH_TIMEOUT_NS) { *a.error = 1; break; }
                }
            }
        }
    }
}

struct FlagArgs {
    int npeers;
    unsigned int epoch[PUSH_MAXPEERS];
    unsigned int *flag[PUSH_MAXPEERS];
    int *error;
};

__global__ void halo_wait_kernel(const FlagArgs a)  // one warp: lane j waits for neighbour j's halo
{
    const int j = threadIdx.x;
    if (j >= a.npeers) return;
    unsigned long long t0 = 0;
    while (ld_acquire_sys(a.flag[j]) < a.epoch[j]) {
        __nanosleep(100);
        const unsigned long long t = push_now();
        if (t0 == 0) t0 = t;
        else if (t - t0 > PUSH_TIMEOUT_NS) { *a.error = 1; return; }
    }
}

__global__ void halo_ack_kernel(const FlagArgs a)  // the ghost entries of this epoch have been consumed
{
    const int j = threadIdx.x;
    if (j < a.npeers) st_release_sys(a.flag[j], a.epoch[j]);
}

static PushVector *push_lookup(nsk_dist_s *D, const double *x)
{
    if (!D->push_ready) return nullptr;
    for (PushVector &v : D->vecs)
        if (v.local == x) return &v;
    return nullptr;
}

// Halo of the first `depth` rings pushed into the neighbours, then the wait for theirs.  Returns NSK_ERR_UNSUPPORTED when x
// is not a registered vector (the caller falls back to NCCL).
static int halo_push(nsk_csr_t A, double *xlocal, int depth)
{
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    PushVector *V = push_lookup(D, xlocal);
    if (!V) return NSK_ERR_UNSUPPORTED;
    if (*D->h_error) {
        nsk_set_error(ctx, "halo push: a bounded wait for a neighbour expired in an earlier call");
        return NSK_ERR_COMM;
    }
    const int slot = (int)(V - D->vecs.data());
    if (V->pending) {
        nsk_set_error(ctx, "halo push: the previous halo of this vector was never released");
        return NSK_ERR_INVALID;
    }
    V->pending = true;
    PushArgs P;
    FlagArgs W;
    P.npeers = W.npeers = 0;
    P.total = 0;
    P.x = xlocal;
    P.begin[0] = 0;
    P.ticket = D->d_ticket;
    P.error = W.error = D->d_error;
    for (size_t pi = 0; pi < D->peers.size(); pi++) {
        DistPeer &Q = D->peers[pi];
        int scnt = 0, rcnt = 0;
        for (int r = 0; r < depth; r++) { scnt += Q.send_ring_count[r]; rcnt += Q.recv_ring_count[r]; }
        if (scnt > 0) {
            const int j = P.npeers++;
            P.send_epoch[j] = ++V->sent[pi];
            P.send_idx[j] = Q.d_send_idx;
            P.dst_idx[j] = Q.d_push_dst;
            P.peer_vec[j] = V->peer[pi];
            P.my_ack[j] = &D->d_flags->ack[slot][pi];
            P.peer_arrive[j] = &Q.peer_flags->arrive[slot][Q.index_at_peer];
            P.total += scnt;
            P.begin[j + 1] = P.total;
        }
        if (rcnt > 0) {
            W.epoch[W.npeers] = ++V->received[pi];
            W.flag[W.npeers++] = &D->d_flags->arrive[slot][pi];
        }
    }
    P.nwait = W.npeers;
    for (int j = 0; j < W.npeers; j++) { P.my_arrive[j] = W.flag[j]; P.wait_epoch[j] = W.epoch[j]; }
    if (P.total > 0) {
        halo_push_kernel<<<(P.total + 255) / 256, 256, 0, ctx->stream>>>(P);
        ctx->launches++;
    } else if (W.npeers > 0) {  // nothing to send, something to receive
        halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(W);
        ctx->launches++;
    }
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}

// After the kernels that read the ghost entries of x: tell the neighbours they may push into x again.
int nsk_halo_release_dev(nsk_csr_t A, const double *xlocal, int depth)
{
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    if (!D) return NSK_OK;
    PushVector *V = push_lookup(D, xlocal);
    if (!V || !V->pending) return NSK_OK;
    V->pending = false;
    const int slot = (int)(V - D->vecs.data());
    FlagArgs K;
    K.npeers = 0;
    K.error = D->d_error;
    for (size_t pi = 0; pi < D->peers.size(); pi++) {
        DistPeer &Q = D->peers[pi];
        int rcnt = 0;
        for (int r = 0; r < depth; r++) rcnt += Q.recv_ring_count[r];
        if (rcnt > 0) {
            K.epoch[K.npeers] = V->received[pi];
            K.flag[K.npeers++] = &Q.peer_flags->ack[slot][Q.index_at_peer];
        }
    }
    if (K.npeers > 0) {
        halo_ack_kernel<<<1, 32, 0, ctx->stream>>>(K);
        ctx->launches++;
        NSK_CUDA(ctx, cudaGetLastError());
    }
    return NSK_OK;
}

// ---- push setup (host layer: handles and layouts travel over torch.distributed / any transport, once) -------------------
// A handle is 80 bytes: the CUDA IPC handle of the ALLOCATION that contains the pointer (cudaMalloc hands out small
// blocks from a shared 2 MB allocation, and an IPC handle always names the whole allocation) + the pointer's offset in it.
NSK_API int nsk_ipc_export(nsk_ctx_t ctx, void *devptr, unsigned char *handle)
{
    if (!ctx || !devptr || !handle) return NSK_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    NSK_CUDA(ctx, cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    NSK_REQUIRE(ctx, fn != nullptr && qr == cudaDriverEntryPointSuccess, "cuMemGetAddressRange is not available");
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<range_fn>(fn)(&base, &size, (unsigned long long)(uintptr_t)devptr) != 0) {
        nsk_set_error(ctx, "nsk_ipc_export: not a device allocation of this process");
        return NSK_ERR_INVALID;
    }
    cudaIpcMemHandle_t h;
    NSK_CUDA(ctx, cudaIpcGetMemHandle(&h, (void *)(uintptr_t)base));
    memset(handle, 0, NSK_IPC_HANDLE_BYTES);
    memcpy(handle, &h, 64);
    const uint64_t off = (uint64_t)((uintptr_t)devptr - (uintptr_t)base);
    memcpy(handle + 64, &off, 8);
    return NSK_OK;
}

// Allocations are mapped once per context (the same allocation may carry several exported pointers) and stay mapped
// until the context is destroyed.
NSK_API int nsk_ipc_import(nsk_ctx_t ctx, const unsigned char *handle, void **peerptr)
{
    if (!ctx || !handle || !peerptr) return NSK_ERR_INVALID;
    NSK_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t off = 0;
    memcpy(&off, handle + 64, 8);
    for (nsk_ipc_mapping &m : ctx->ipc_maps)
        if (!memcmp(m.handle, handle, 64)) {
            *peerptr = (char *)m.base + off;
            return NSK_OK;
        }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *base = nullptr;
    NSK_CUDA(ctx, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    nsk_ipc_mapping m;
    memcpy(m.handle, handle, 64);
    m.base = base;
    ctx->ipc_maps.push_back(m);
    *peerptr = (char *)base + off;
    return NSK_OK;
}

void nsk_ipc_close_all(nsk_ctx_t ctx)
{
    for (nsk_ipc_mapping &m : ctx->ipc_maps) cudaIpcCloseMemHandle(m.base);
    ctx->ipc_maps.clear();
}

// This rank's flag block (allocated on first use), to be exported to the neighbours.
NSK_API int nsk_dist_push_flags(nsk_csr_t A, void **flags)
{
    if (!A || !A->dist || !flags) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    NSK_REQUIRE(ctx, (int)D->peers.size() <= PUSH_MAXPEERS, "halo push supports up to 8 neighbours");
    if (!D->d_flags) {
        NSK_CUDA(ctx, cudaMalloc(&D->d_flags, sizeof(PushFlags)));
        NSK_CUDA(ctx, cudaMemset(D->d_flags, 0, sizeof(PushFlags)));
        NSK_CUDA(ctx, cudaMalloc(&D->d_ticket, sizeof(unsigned int)));
        NSK_CUDA(ctx, cudaMemset(D->d_ticket, 0, sizeof(unsigned int)));
        NSK_CUDA(ctx, cudaHostAlloc(&D->h_error, sizeof(int), cudaHostAllocMapped));
        *D->h_error = 0;
        NSK_CUDA(ctx, cudaHostGetDevicePointer(&D->d_error, D->h_error, 0));
    }
    *flags = D->d_flags;
    return NSK_OK;
}

// Where the entries this rank RECEIVES from `peer_rank` land in its local vector, ring by ring (the sender needs it).
NSK_API int nsk_dist_recv_layout(nsk_csr_t A, int peer_rank, int *ring_start, int *ring_count)
{
    if (!A || !A->dist || !ring_start || !ring_count) return NSK_ERR_INVALID;
    for (DistPeer &Q : A->dist->peers)
        if (Q.rank == peer_rank) {
            for (int r = 0; r < A->dist->depth; r++) { ring_start[r] = Q.recv_ring_start[r]; ring_count[r] = Q.recv_ring_count[r]; }
            return NSK_OK;
        }
    return NSK_ERR_INVALID;
}

NSK_API int nsk_dist_peer_count(nsk_csr_t A) { return A && A->dist ? (int)A->dist->peers.size() : 0; }
NSK_API int nsk_dist_peer_rank(nsk_csr_t A, int index)
{
    return A && A->dist && index >= 0 && index < (int)A->dist->peers.size() ? A->dist->peers[(size_t)index].rank : -1;
}

// What the neighbour `peer_rank` told us: our position in ITS peer list, where our entries land in ITS local vector
// (its recv layout for us), and its flag block mapped here (nsk_ipc_import).
NSK_API int nsk_dist_push_peer(nsk_csr_t A, int peer_rank, int index_at_peer, const int *peer_ring_start, const int *peer_ring_count,
                               void *peer_flags)
{
    if (!A || !A->dist || !peer_ring_start || !peer_ring_count || !peer_flags) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    NSK_REQUIRE(ctx, index_at_peer >= 0 && index_at_peer < PUSH_MAXPEERS, "peer index out of range");
    for (DistPeer &Q : D->peers)
        if (Q.rank == peer_rank) {
            std::vector<int> dst;
            for (int r = 0; r < D->depth; r++) {
                NSK_REQUIRE(ctx, peer_ring_count[r] == Q.send_ring_count[r], "the neighbour expects a different halo size");
                for (int q = 0; q < Q.send_ring_count[r]; q++) dst.push_back(peer_ring_start[r] + q);
            }
            if (!dst.empty()) {
                if (Q.d_push_dst) cudaFree(Q.d_push_dst);
                NSK_CUDA(ctx, cudaMalloc(&Q.d_push_dst, sizeof(int) * dst.size()));
                NSK_CUDA(ctx, cudaMemcpy(Q.d_push_dst, dst.data(), sizeof(int) * dst.size(), cudaMemcpyHostToDevice));
            }
            Q.index_at_peer = index_at_peer;
            Q.peer_flags = static_cast<PushFlags *>(peer_flags);
            bool all = D->d_flags != nullptr;
            for (DistPeer &R : D->peers) all = all && R.peer_flags != nullptr;
            D->push_ready = all;
            return NSK_OK;
        }
    nsk_set_error(ctx, "rank %d is not a neighbour of this operator", peer_rank);
    return NSK_ERR_INVALID;
}

// Registers a local vector (n_cols_local doubles, allocated with nsk_malloc) together with the neighbours' copies of
// the SAME vector mapped here (peer_ptrs[i] belongs to neighbour i of nsk_dist_peer_rank).  Matrix-powers calls whose x is
// a registered vector exchange the halo by pushing; all ranks must register the same vectors in the same order.
NSK_API int nsk_dist_vector_register(nsk_csr_t A, double *local, double *const *peer_ptrs)
{
    if (!A || !A->dist || !local || !peer_ptrs) return NSK_ERR_INVALID;
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    NSK_REQUIRE(ctx, D->push_ready, "halo push is not set up (nsk_dist_push_flags / nsk_dist_push_peer for every neighbour)");
    NSK_REQUIRE(ctx, (int)D->vecs.size() < PUSH_SLOTS, "too many registered vectors");
    PushVector V;
    V.local = local;
    for (size_t i = 0; i < D->peers.size(); i++) V.peer[i] = peer_ptrs[i];
    D->vecs.push_back(V);
    return NSK_OK;
}

// allow_push: the caller promises to call nsk_halo_release_dev(A, xlocal, depth) after the kernels that read the ghost
// entries (only then may a registered vector take the push path).
int nsk_halo_exchange_dev(nsk_csr_t A, double *xlocal, int depth, bool allow_push)
{
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    if (!D) return NSK_OK;
    NSK_REQUIRE(ctx, depth >= 1 && depth <= D->depth, "halo depth exceeds the plan's depth");
    if (D->peers.empty()) return NSK_OK;
    if (ctx->opt.halo_push < 0) return NSK_OK;  // MEASUREMENT ONLY (tools/dist_probe.py): no exchange, ghost entries stay as they are
    if (allow_push && ctx->opt.halo_push) {  // registered vectors: one push kernel over NVLink peer memory instead of pack + NCCL + unpack
        const int s = halo_push(A, xlocal, depth);
        if (s != NSK_ERR_UNSUPPORTED) return s;
    }
    NSK_REQUIRE(ctx, nsk_comm_active(ctx), "the operator has peers but no communicator is attached (nsk_comm_init)");
    // One message per peer and direction, ALWAYS (the protocol must not depend on anything only this rank knows, such
    // as its own number of peers: both ends of a message have to agree on its size): the first `depth` rings of a
    // peer's send list are contiguous in its send buffer (ring-major), and land contiguously in d_recvbuf; a ring's
    // slice from one peer is contiguous in the local vector, so unpacking is `depth` straight copies per peer.  Pack
    // and unpack run as one launch per HALO_MAX_SEG segments (one launch each up to 64 peers x rings).
    struct Seg { const double *src; double *dst; const int *idx; int count; };
    std::vector<Seg> pack, unpack;
    std::vector<const double *> sendbuf;
    std::vector<int> sendcount, recvcount, peer;
    std::vector<double *> recvbuf;
    for (DistPeer &P : D->peers) {
        int scnt = 0, rcnt = 0;
        for (int r = 0; r < depth; r++) { scnt += P.send_ring_count[r]; rcnt += P.recv_ring_count[r]; }
        if (scnt > 0) pack.push_back(Seg{xlocal, P.d_sendbuf, P.d_send_idx, scnt});
        int off = 0;
        for (int r = 0; r < depth && depth > 1; r++) {  // depth 1: the single ring is received in place, nothing to unpack
            if (P.recv_ring_count[r] == 0) continue;
            unpack.push_back(Seg{P.d_recvbuf + off, xlocal + P.recv_ring_start[r], nullptr, P.recv_ring_count[r]});
            off += P.recv_ring_count[r];
        }
        peer.push_back(P.rank);
        sendbuf.push_back(P.d_sendbuf);
        sendcount.push_back(scnt);
        recvbuf.push_back(depth > 1 ? P.d_recvbuf : xlocal + P.recv_ring_start[0]);
        recvcount.push_back(rcnt);
    }
    auto move = [&](const std::vector<Seg> &segs) {
        for (size_t s0 = 0; s0 < segs.size(); s0 += HALO_MAX_SEG) {
            HaloSegs H;
            H.nseg = 0;
            H.total = 0;
            H.begin[0] = 0;
            for (size_t i = s0; i < std::min(segs.size(), s0 + (size_t)HALO_MAX_SEG); i++) {
                H.src[H.nseg] = segs[i].src;
                H.dst[H.nseg] = segs[i].dst;
                H.idx[H.nseg] = segs[i].idx;
                H.total += segs[i].count;
                H.begin[++H.nseg] = H.total;
            }
            if (H.total > 0) {
                halo_move_kernel<<<(H.total + 255) / 256, 256, 0, ctx->stream>>>(H);
                ctx->launches++;
            }
        }
    };
    move(pack);
    NSK_CUDA(ctx, cudaGetLastError());
    NSK_TRY(nsk_comm_sendrecv(ctx, (int)peer.size(), peer.data(), sendbuf.data(), sendcount.data(), recvbuf.data(),
                              recvcount.data()));
    move(unpack);
    NSK_CUDA(ctx, cudaGetLastError());
    return NSK_OK;
}

NSK_API int nsk_halo_exchange(nsk_csr_t A, double *xlocal, int depth)
{
    if (!A || !xlocal) return NSK_ERR_INVALID;
    if (!A->dist) {
        nsk_set_error(A->ctx, "nsk_halo_exchange: not a distributed operator");
        return NSK_ERR_INVALID;
    }
    NSK_CUDA(A->ctx, cudaSetDevice(A->ctx->device));
    NSK_TRY(nsk_halo_exchange_dev(A, xlocal, depth, true));
    return nsk_halo_release_dev(A, xlocal, depth);  // registered vector: the ghost entries are the caller's from here on
}

// levels[l] are LOCAL vectors (n_cols_local doubles); level l is valid on its row prefix, the owned part
// of every level is the distributed result.  d_x is a local vector whose owned part is filled.
int nsk_dist_mpk(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode)
{
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    NSK_REQUIRE(ctx, k <= D->depth, "k exceeds the halo depth the operator was planned for");
    NSK_TRY(nsk_halo_exchange_dev(A, const_cast<double *>(d_x), k, true));
    int level_rows[NSK_MAX_K];
    for (int l = 0; l < k; l++) level_rows[l] = D->ring_start[k - l];
    NSK_TRY(nsk_mpk_local(A, k, d_x, d_levels, mode, level_rows));
    return nsk_halo_release_dev(A, d_x, k);
}

// Two right-hand sides: both halos first (two exchanges back to back on the stream), then one fused sweep for both.
int nsk_dist_mpk2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                  double *const *d_levels2, nsk_mode mode)
{
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    NSK_REQUIRE(ctx, k <= D->depth, "k exceeds the halo depth the operator was planned for");
    NSK_TRY(nsk_halo_exchange_dev(A, const_cast<double *>(d_x), k, true));
    NSK_TRY(nsk_halo_exchange_dev(A, const_cast<double *>(d_x2), k, true));
    int level_rows[NSK_MAX_K];
    for (int l = 0; l < k; l++) level_rows[l] = D->ring_start[k - l];
    NSK_TRY(nsk_mpk_local2(A, k, d_x, d_levels, d_x2, d_levels2, mode, level_rows));
    NSK_TRY(nsk_halo_release_dev(A, d_x, k));
    return nsk_halo_release_dev(A, d_x2, k);
}
