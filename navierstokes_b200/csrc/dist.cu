// dist.cu -- device side of the row-partitioned operator: halo exchange (pack kernel + grouped
// NCCL point-to-point over NVLink) and matrix powers with redundant ghost levels.
//
// The reference is single-process; this is the multi-GPU layer BASELINE.json's north
// star asks for: one process per GPU, the depth-k halo exchanged ONCE per matrix-powers call, ghost
// levels recomputed redundantly on shrinking row prefixes (plan: dist_plan.cpp).
#include <algorithm>

#include "dist_plan.h"
#include "nsk_internal.h"

int nsk_mpk_local(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode,
                  const int *level_rows);  // mpk.cu
int nsk_mpk_local2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                   double *const *d_levels2, nsk_mode mode, const int *level_rows);  // mpk.cu
void nsk_wave_set_block_extents(nsk_csr_t A, const int *ptrow, const int *indcol);
int nsk_comm_rank(nsk_ctx_t ctx);
int nsk_comm_size(nsk_ctx_t ctx);

struct DistPeer {
    int rank = 0;
    int *d_send_idx = nullptr;          // local indices to pack, ring-major
    double *d_sendbuf = nullptr;
    double *d_recvbuf = nullptr;        // one contiguous message per peer lands here, then unpack_kernel scatters it
    std::vector<int> send_ring_count;   // depth
    std::vector<int> recv_ring_count;   // depth
    std::vector<int> recv_ring_start;   // depth, offset into the local vector
};

struct nsk_dist_s {
    int n_owned = 0, depth = 0;
    std::vector<int> ring_start;  // depth + 2
    std::vector<DistPeer> peers;
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
    }
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

int nsk_halo_exchange_dev(nsk_csr_t A, double *xlocal, int depth)
{
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    if (!D) return NSK_OK;
    NSK_REQUIRE(ctx, depth >= 1 && depth <= D->depth, "halo depth exceeds the plan's depth");
    if (D->peers.empty()) return NSK_OK;
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
    return nsk_halo_exchange_dev(A, xlocal, depth);
}

// levels[l] are LOCAL vectors (n_cols_local doubles); level l is valid on its row prefix, the owned part
// of every level is the distributed result.  d_x is a local vector whose owned part is filled.
int nsk_dist_mpk(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, nsk_mode mode)
{
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    NSK_REQUIRE(ctx, k <= D->depth, "k exceeds the halo depth the operator was planned for");
    NSK_TRY(nsk_halo_exchange_dev(A, const_cast<double *>(d_x), k));
    int level_rows[NSK_MAX_K];
    for (int l = 0; l < k; l++) level_rows[l] = D->ring_start[k - l];
    return nsk_mpk_local(A, k, d_x, d_levels, mode, level_rows);
}

// Two right-hand sides: both halos first (two exchanges back to back on the stream), then one fused sweep for both.
int nsk_dist_mpk2(nsk_csr_t A, int k, const double *d_x, double *const *d_levels, const double *d_x2,
                  double *const *d_levels2, nsk_mode mode)
{
    nsk_ctx_t ctx = A->ctx;
    nsk_dist_s *D = A->dist;
    NSK_REQUIRE(ctx, k <= D->depth, "k exceeds the halo depth the operator was planned for");
    NSK_TRY(nsk_halo_exchange_dev(A, const_cast<double *>(d_x), k));
    NSK_TRY(nsk_halo_exchange_dev(A, const_cast<double *>(d_x2), k));
    int level_rows[NSK_MAX_K];
    for (int l = 0; l < k; l++) level_rows[l] = D->ring_start[k - l];
    return nsk_mpk_local2(A, k, d_x, d_levels, d_x2, d_levels2, mode, level_rows);
}
