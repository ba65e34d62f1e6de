"""CPU emulation of the distributed matrix-powers path for the gloo tests.

The PLAN (ghost rings, local numbering, per-peer lists) comes from the product's host-only planner
(nsk_plan_* in libnsk.so); the data motion is done with torch.distributed (gloo) point-to-point exactly
as dist.cu does it with NCCL (one message per peer and ring, received straight into the local vector);
the arithmetic is the CPU oracle applied to the level prefixes.  Test infrastructure only.
"""
import numpy as np
import torch
import torch.distributed as dist


def halo_exchange(plan, xlocal: np.ndarray, depth: int):
    """Refresh rings 1..depth of xlocal in place."""
    ops, keep = [], []
    ghosts = plan.ghosts()
    # sends: what each peer asked of us, ring-major, first `depth` rings
    for peer, (gids, rc) in sorted(plan.sends.items()):
        off = 0
        for r in range(depth):
            c = int(rc[r])
            if c:
                t = torch.from_numpy(xlocal[gids[off:off + c] - plan.row_starts[plan.rank]].copy())
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, peer))
            off += c
    recvs = []
    for peer, (gids, rc) in sorted(plan.my_requests.items()):
        off = 0
        for r in range(depth):
            c = int(rc[r])
            if c:
                ring = ghosts[plan.ring_start[r + 1] - plan.n_owned: plan.ring_start[r + 2] - plan.n_owned]
                start = int(plan.ring_start[r + 1] + np.searchsorted(ring, gids[off]))
                t = torch.empty(c, dtype=torch.float64)
                recvs.append((start, c, t))
                ops.append(dist.P2POp(dist.irecv, t, peer))
            off += c
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for start, c, t in recvs:
        xlocal[start:start + c] = t.numpy()


def mpk(plan, oracle_lib, x_owned: np.ndarray, k: int):
    """Owned parts of A^1..k x computed the distributed way (one depth-k exchange, redundant ghosts)."""
    lc = plan.local_csr()
    src = np.zeros(plan.n_cols_local)
    src[:plan.n_owned] = x_owned
    halo_exchange(plan, src, k)
    out = []
    for l in range(k):
        rows = int(plan.ring_start[k - l])
        y = np.zeros(plan.n_cols_local)
        nz = lc.ptrow[rows]
        y[:rows] = oracle_lib.spmv(lc.ptrow[:rows + 1], lc.indcol[:nz], lc.coef[:nz], src)
        out.append(y[:plan.n_owned].copy())
        src = y
    return out


def cg(plan, oracle_lib, b_owned: np.ndarray, tol: float, maxit: int):
    """Textbook CG with one depth-1 exchange per product and all-reduced dots (what cg.cu does)."""
    lc = plan.local_csr()
    n = plan.n_owned
    nz = lc.ptrow[n]

    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    x = np.zeros(n)
    r = b_owned.copy()
    p = np.zeros(plan.n_cols_local)
    p[:n] = r
    bb = allsum(float(r @ r))
    rr = bb
    it = 0
    while it < maxit and np.sqrt(rr / bb) > tol:
        halo_exchange(plan, p, 1)
        q = oracle_lib.spmv(lc.ptrow[:n + 1], lc.indcol[:nz], lc.coef[:nz], p)
        alpha = rr / allsum(float(p[:n] @ q))
        x += alpha * p[:n]
        r -= alpha * q
        rr_new = allsum(float(r @ r))
        p[:n] = r + (rr_new / rr) * p[:n]
        rr = rr_new
        it += 1
    return x, it, float(np.sqrt(rr / bb))
