"""CPU prototype (research record, not product code): aggregation multigrid as the PCG
preconditioner for K_ff of the synthetic occupancy grids.  Node-based aggregates (3 DOF per
aggregate, piecewise-constant prolongation per component: the operator's near-null space is the
three translations -- the transverse spring resists rotations), Galerkin coarse operators,
3x3-block-Jacobi / Chebyshev smoothing, V / W cycles.

    python tools/proto/amg_proto.py N [options]
"""
import argparse
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, ".")
from oracle import fea_oracle as fo
from mycelium_fea_project_b200.synth import synth_network


def build_problem(N, case="Y"):
    coords, n1, n2 = synth_network(N)
    K = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
    axis, comp = {"Y": (1, 1), "X": (0, 0)}[case]
    hi, lo = fo.grip_nodes(coords, 1.5, axis)
    kd, kv = fo.build_bc(hi, lo, 0.02, -0.02, comp)
    free, K_ff, F_f = fo.reduce_system(K, kd, kv)
    assert len(free) % 3 == 0
    return sp.bsr_matrix(K_ff, blocksize=(3, 3)), F_f, coords, free


def block_diag_inv(A):
    """inverse of the 3x3 diagonal blocks of a BSR matrix -> (n,3,3)"""
    n = A.shape[0] // 3
    D = np.zeros((n, 3, 3))
    indptr, indices, data = A.indptr, A.indices, A.data
    rows = np.repeat(np.arange(n), np.diff(indptr))
    m = indices == rows
    D[rows[m]] = data[m]
    # guard singular blocks
    for i in np.where(np.abs(np.linalg.det(D)) < 1e-300)[0]:
        D[i] += np.eye(3) * 1e-12
    return np.linalg.inv(D)


def node_strength(A):
    """scalar node graph: w_ij = -trace(K_ij) (sum of spring stiffnesses), i != j"""
    n = A.shape[0] // 3
    indptr, indices, data = A.indptr, A.indices, A.data
    rows = np.repeat(np.arange(n), np.diff(indptr))
    w = -np.trace(data, axis1=1, axis2=2)
    m = indices != rows
    return sp.csr_matrix((w[m], (rows[m], indices[m])), shape=(n, n))


def pairwise_match(W, rounds=4, state=None):
    """handshake matching: every unmatched node points at its strongest unmatched neighbour;
    mutual pointers become a pair.  Returns agg id per node (pairs and leftover singletons)."""
    n = W.shape[0]
    W = W.tocsr()
    matched = np.full(n, -1, dtype=np.int64)
    indptr, indices, w = W.indptr, W.indices, W.data
    rows = np.repeat(np.arange(n), np.diff(indptr))
    for _ in range(rounds):
        ok = (matched[rows] < 0) & (matched[indices] < 0) & (w > 0)
        if not ok.any():
            break
        r, c, ww = rows[ok], indices[ok], w[ok]
        # strongest neighbour per row (ties -> smallest column)
        order = np.lexsort((c, -ww, r))
        r_s, c_s = r[order], c[order]
        first = np.ones(len(r_s), bool)
        first[1:] = r_s[1:] != r_s[:-1]
        best = np.full(n, -1, dtype=np.int64)
        best[r_s[first]] = c_s[first]
        i = np.where(best >= 0)[0]
        mutual = i[best[best[i]] == i]
        matched[mutual] = best[mutual]
    agg = np.full(n, -1, dtype=np.int64)
    a = 0
    lead = np.where((matched >= 0) & (np.arange(n) < matched))[0]
    agg[lead] = np.arange(len(lead))
    agg[matched[lead]] = agg[lead]
    a = len(lead)
    single = np.where(agg < 0)[0]
    agg[single] = a + np.arange(len(single))
    return agg, a + len(single)


def join_singletons(W, agg, n_agg):
    """attach leftover singletons that still have a neighbour to the neighbour's aggregate
    (strongest), so aggregates have 2-3 nodes"""
    n = W.shape[0]
    size = np.bincount(agg, minlength=n_agg)
    W = W.tocsr()
    indptr, indices, w = W.indptr, W.indices, W.data
    rows = np.repeat(np.arange(n), np.diff(indptr))
    sing = size[agg] == 1
    ok = sing[rows] & ~sing[indices] & (w > 0)
    r, c, ww = rows[ok], indices[ok], w[ok]
    order = np.lexsort((c, -ww, r))
    r_s, c_s = r[order], c[order]
    first = np.ones(len(r_s), bool)
    first[1:] = r_s[1:] != r_s[:-1]
    agg = agg.copy()
    agg[r_s[first]] = agg[c_s[first]]
    # renumber
    u, inv = np.unique(agg, return_inverse=True)
    return inv, len(u)


def aggregate(A, passes=2, join=True):
    """passes of pairwise aggregation on the node graph -> agg id per fine node"""
    n = A.shape[0] // 3
    agg_total = np.arange(n)
    W = node_strength(A)
    n_cur = n
    for p in range(passes):
        agg, n_agg = pairwise_match(W)
        if join:
            agg, n_agg = join_singletons(W, agg, n_agg)
        agg_total = agg[agg_total]
        P = sp.csr_matrix((np.ones(n_cur), (np.arange(n_cur), agg)), shape=(n_cur, n_agg))
        W = (P.T @ W @ P).tocsr()
        W.setdiag(0)
        W.eliminate_zeros()
        n_cur = n_agg
    return agg_total, n_cur


def mis_aggregate(A, theta=0.0, prio="hash"):
    """root-point aggregation: roots = maximal independent set of the node graph (Luby rounds with a
    fixed hash priority), every other node joins its most strongly coupled root."""
    n = A.shape[0] // 3
    W = node_strength(A).tocsr()
    indptr, indices, w = W.indptr, W.indices, W.data
    rows = np.repeat(np.arange(n), np.diff(indptr))
    if theta > 0:
        wmax = np.zeros(n)
        np.maximum.at(wmax, rows, w)
        strong = w >= theta * wmax[rows]
    else:
        strong = w > 0
    r_, c_, w_ = rows[strong], indices[strong], w[strong]
    if prio == "hash":
        pri = ((np.arange(n, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(2**32)).astype(np.float64)
    else:
        deg = np.bincount(r_, minlength=n).astype(np.float64)
        pri = deg * 2**32 + ((np.arange(n, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(2**32)).astype(np.float64)
    state = np.zeros(n, dtype=np.int8)      # 0 undecided 1 root 2 covered
    rounds = 0
    while (state == 0).any():
        rounds += 1
        und = state == 0
        e = und[r_] & und[c_]
        nbmax = np.full(n, -1.0)
        np.maximum.at(nbmax, r_[e], pri[c_[e]])
        new_root = und & (pri > nbmax)
        state[new_root] = 1
        cov = np.zeros(n, bool)
        cov[r_[new_root[c_]]] = True
        state[cov & (state == 0)] = 2
    roots = state == 1
    rid = np.cumsum(roots) - 1
    agg = np.where(roots, rid, -1)
    # covered nodes: strongest root neighbour
    e = roots[c_] & ~roots[r_]
    rr, cc, ww = r_[e], c_[e], w_[e]
    order = np.lexsort((cc, -ww, rr))
    rs, cs = rr[order], cc[order]
    first = np.ones(len(rs), bool)
    first[1:] = rs[1:] != rs[:-1]
    agg[rs[first]] = rid[cs[first]]
    assert (agg >= 0).all()
    return agg, int(roots.sum()), rounds


class Level:
    pass


def build_hierarchy(A, passes=2, max_levels=12, min_nodes=200, drop_isolated=True, verbose=True):
    levels = []
    while True:
        L = Level()
        L.A = A.tocsr()
        L.Ab = A
        L.Dinv = block_diag_inv(A)
        L.n = A.shape[0] // 3
        levels.append(L)
        if len(levels) >= max_levels or L.n <= min_nodes:
            break
        if MIS:
            agg, n_agg, rounds = mis_aggregate(A, THETA, PRIO)
            if verbose: print("  MIS rounds", rounds, "n", L.n, "->", n_agg)
        else:
            agg, n_agg = aggregate(A, passes)
        if drop_isolated:
            # aggregates without any outside coupling need no coarse representation
            Pn = sp.csr_matrix((np.ones(L.n), (np.arange(L.n), agg)), shape=(L.n, n_agg))
            Wc = (Pn.T @ node_strength(A) @ Pn).tocsr()
            Wc.setdiag(0)
            Wc.eliminate_zeros()
            keep = np.diff(Wc.indptr) > 0
            newid = np.cumsum(keep) - 1
            agg = np.where(keep[agg], newid[agg], -1)
            n_agg = int(keep.sum())
        if n_agg == 0 or n_agg >= 0.9 * L.n:
            break
        m = agg >= 0
        rows = (3 * np.where(m)[0][:, None] + np.arange(3)).ravel()
        cols = (3 * agg[m][:, None] + np.arange(3)).ravel()
        L.P = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(3 * L.n, 3 * n_agg))
        Ac = (L.P.T @ L.A @ L.P).tocsr()
        A = sp.bsr_matrix(Ac, blocksize=(3, 3))
    if verbose:
        print("levels:", [(l.n, l.Ab.nnz // 9) for l in levels], "op complexity %.2f" % (sum(l.Ab.nnz for l in levels) / levels[0].Ab.nnz))
    return levels


def apply_dinv(L, r):
    return np.einsum("nij,nj->ni", L.Dinv, r.reshape(-1, 3)).ravel()


def smooth(L, x, b, nu, omega, cheb=None):
    if cheb:
        # Chebyshev polynomial smoother on D^-1 A, eigenvalue interval [lmax/alpha, lmax]
        lmax, alpha, deg = cheb
        lmin = lmax / alpha
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho = 1.0 / sigma
        r = b - L.A @ x if x.any() else b.copy()
        d = apply_dinv(L, r) / theta
        x = x + d
        for _ in range(deg - 1):
            r = r - L.A @ d
            rho_new = 1.0 / (2 * sigma - rho)
            d = rho_new * rho * d + (2 * rho_new / delta) * apply_dinv(L, r)
            x = x + d
            rho = rho_new
        return x
    for _ in range(nu):
        r = b - L.A @ x if x.any() else b
        x = x + omega * apply_dinv(L, r)
    return x


def est_lmax(L, iters=15):
    rng = np.random.default_rng(1)
    v = rng.standard_normal(L.A.shape[0])
    lam = 1.0
    for _ in range(iters):
        w = apply_dinv(L, L.A @ v)
        lam = np.linalg.norm(w) / np.linalg.norm(v)
        v = w / np.linalg.norm(w)
    return lam


def cycle(levels, l, b, opt):
    L = levels[l]
    if l == len(levels) - 1:
        x = np.zeros_like(b)
        return smooth(L, x, b, opt.coarse_sweeps, opt.omega, L.cheb if opt.cheb else None)
    x = smooth(L, np.zeros_like(b), b, opt.nu, opt.omega, L.cheb if opt.cheb else None)
    for k in range(opt.gamma if l >= opt.wfrom else 1):
        r = b - L.A @ x
        ec = cycle(levels, l + 1, L.P.T @ r, opt)
        x = x + opt.scale * (L.P @ ec)
    x = smooth(L, x, b, opt.nu, opt.omega, L.cheb if opt.cheb else None)
    return x


def pcg(A, b, M, rtol=1e-10, maxit=20000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bb = np.sqrt(b @ b)
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.sqrt(r @ r) <= rtol * bb:
            return x, it
        z = M(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit


MIS = False
THETA = 0.0
PRIO = "hash"


def main():
    global MIS, THETA, PRIO
    ap = argparse.ArgumentParser()
    ap.add_argument("--mis", action="store_true")
    ap.add_argument("--theta", type=float, default=0.0)
    ap.add_argument("--prio", default="hash")
    ap.add_argument("N", type=int)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--nu", type=int, default=1)
    ap.add_argument("--omega", type=float, default=0.7)
    ap.add_argument("--gamma", type=int, default=1)
    ap.add_argument("--wfrom", type=int, default=0)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cheb", type=int, default=0, help="Chebyshev degree (0 = damped block Jacobi)")
    ap.add_argument("--cheb-alpha", type=float, default=8.0)
    ap.add_argument("--coarse-sweeps", type=int, default=8)
    ap.add_argument("--min-nodes", type=int, default=200)
    ap.add_argument("--max-levels", type=int, default=12)
    ap.add_argument("--case", default="Y")
    ap.add_argument("--baseline", action="store_true")
    opt = ap.parse_args()
    MIS, THETA, PRIO = opt.mis, opt.theta, opt.prio
    A, b, coords, free = build_problem(opt.N, opt.case)
    print("n_free", A.shape[0], "blocks", A.nnz // 9)
    t0 = time.time()
    levels = build_hierarchy(A, opt.passes, opt.max_levels, opt.min_nodes)
    for L in levels:
        L.cheb = (1.1 * est_lmax(L), opt.cheb_alpha, opt.cheb) if opt.cheb else None
    print("setup %.1fs" % (time.time() - t0))
    Acsr = A.tocsr()
    if opt.baseline:
        x, it = pcg(Acsr, b, lambda r: apply_dinv(levels[0], r))
        print("block3-Jacobi PCG iterations", it)
    t0 = time.time()
    x, it = pcg(Acsr, b, lambda r: cycle(levels, 0, r, opt))
    res = np.linalg.norm(b - Acsr @ x) / np.linalg.norm(b)
    # work per iteration in fine-SpMV equivalents
    nnz = [l.Ab.nnz for l in levels]
    visits = [(opt.gamma ** max(0, i - opt.wfrom)) if opt.gamma > 1 else 1 for i in range(len(levels))]
    sm = (opt.cheb if opt.cheb else opt.nu)
    work = 1 + sum(v * n * (2 * sm + 1) for v, n in zip(visits, nnz)) / nnz[0]
    print(f"AMG-PCG iterations {it}  true relres {res:.2e}  time {time.time() - t0:.1f}s  ~work/iter {work:.1f} fine SpMV -> {it * work:.0f} SpMV-equivalents")


if __name__ == "__main__":
    main()
