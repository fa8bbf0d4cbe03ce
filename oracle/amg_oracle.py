"""numpy restatement of the GPU path's aggregation-multigrid preconditioner -- TEST INFRASTRUCTURE ONLY.

The reference solves K_ff U_f = F_f directly (src/fea_solver.py:128, SuperLU) and its PETSc variants
offer {jacobi, sor, ilu, icc, gamg} (src/fea_petsc_solverAndPC.cpp:330-331).  The CUDA path's default
preconditioner (csrc/amg_setup.cu, csrc/pcg_amg.cu) is an unsmoothed-aggregation multigrid V-cycle on
the 3x3 node-block operator -- the GAMG idea restated for this operator, whose near-null space is the
three translations (the transverse spring resists rotation).  This file states the SAME algorithm,
step for step and with the same tie rules, in numpy, so that tests can check the GPU's aggregates,
level sizes and iteration counts against it ("iteration count equal to a numpy PCG with the same
preconditioner").  There is nothing in the reference to pin it to: its pin is the reference's direct
solve (tests compare U of an AMG-PCG solve with fea_oracle.solve_system) -- "parity unpinned" as an
algorithm, pinned as a result.

Algorithm per level (node = 3 DOF, blocks = 3x3, symmetric):
  strength   w_ij = -((xx + yy) + zz) of block (i, j), i != j, j active and rank-local; only w > 0 counts
  propose    best[i] = neighbour of largest w (ties: smaller j)
  accept     i, j are a pair iff best[i] == j and best[j] == i
  join       an unpaired node joins the pair of its strongest PAIRED neighbour (ties: smaller j)
  keep       aggregates without any block to an active node outside themselves are not represented
             on the next level (floating pieces that have collapsed to one aggregate)
  number     kept aggregates in the order of their smallest member ("root")
  coarse     A_c[I, J] = sum of the fine blocks (i, j), i in I, j in J, added in fine block order
  smoother   damped 3x3-block Jacobi, e += omega * D^-1 (r - A e), D = diag block + reg I
  cycle      V(1,1): pre-smooth from zero, restrict the residual, recurse, add SCALE * correction,
             post-smooth; the coarsest level does COARSE_SWEEPS smoother sweeps
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

OMEGA = 0.9            # smoother damping
SCALE = 1.5            # over-correction of the piecewise-constant coarse correction
COARSE_SWEEPS = 8
MIN_NODES = 200        # a level with at most this many nodes is the coarsest
MAX_LEVELS = 16
MAX_RATIO = 0.8        # stop if a level does not shrink below MAX_RATIO * n
PC_FP32 = True         # the V-cycle's level operators are stored rounded to FP32 (csrc/amg_sweep.cuh); products and sums
                       # stay FP64, the CG operator stays FP64.  False restates MYC_AMG_FP64=1.
REPLICATE_NODES = 65536  # multi-GPU: a level with at most this many nodes (over all ranks) is held by every rank
                         # in full, so from there on aggregates are formed without regard to the row partition


class Level:
    """n, brp (n+1), bnode (nb), bval (nb,6: xx xy xz yy yz zz), act (n) bool, dinv (n,6), agg (n), n_coarse"""


def level_from_csr(K, free_mask, reg=1e-12):
    """Level 0 from a scalar CSR with the 3x3 node-block structure; free_mask (n_dof) bool.
    Returns None when a node is only partially free (the AMG path needs node-complete Dirichlet sets)."""
    K = sp.csr_matrix(K)
    n = K.shape[0] // 3
    fm = np.asarray(free_mask, bool).reshape(n, 3)
    if np.any(fm.any(1) != fm.all(1)):
        return None
    B = sp.bsr_matrix(K, blocksize=(3, 3))
    B.sort_indices()
    L = Level()
    L.n = n
    L.brp = B.indptr.astype(np.int64)
    L.bnode = B.indices.astype(np.int64)
    d = B.data
    L.bval = np.stack([d[:, 0, 0], d[:, 0, 1], d[:, 0, 2], d[:, 1, 1], d[:, 1, 2], d[:, 2, 2]], axis=1)
    L.act = fm.all(1)
    L.reg = reg
    return L


def _rows(L):
    return np.repeat(np.arange(L.n), np.diff(L.brp))


def _strongest(rows, cols, w, n):
    """per row: column of the largest w (ties: smaller column); -1 for rows without candidates"""
    best = np.full(n, -1, dtype=np.int64)
    if len(rows):
        order = np.lexsort((cols, -w, rows))
        r, c = rows[order], cols[order]
        first = np.ones(len(r), bool)
        first[1:] = r[1:] != r[:-1]
        best[r[first]] = c[first]
    return best


def aggregate(L):
    """agg (n,) coarse id or -1, n_coarse; follows csrc/amg_setup.cu kernel by kernel"""
    n = L.n
    rows, cols = _rows(L), L.bnode
    w = -((L.bval[:, 0] + L.bval[:, 3]) + L.bval[:, 5])
    off = (cols != rows) & L.act[rows] & L.act[cols]
    cand = off & (w > 0)
    owner = getattr(L, "owner", None)
    if owner is not None:                     # row-partitioned hierarchy: aggregates never span ranks
        cand &= owner[rows] == owner[cols]
    best = _strongest(rows[cand], cols[cand], w[cand], n)                       # propose
    idx = np.arange(n)
    paired = np.where((best >= 0) & (best[np.maximum(best, 0)] == idx), best, -1)   # accept
    single = (paired < 0) & L.act
    cj = cand & single[rows] & (paired[cols] >= 0)
    join_to = _strongest(rows[cj], cols[cj], w[cj], n)                          # join
    root = np.where(paired >= 0, np.minimum(idx, paired), idx)
    j = join_to >= 0
    root[j] = np.minimum(join_to[j], paired[join_to[j]])
    keep = np.zeros(n, bool)
    ext = off & (root[rows] != root[cols])
    keep[root[rows[ext]]] = True                                                # keep
    lead = (root == idx) & L.act & keep
    cid = np.cumsum(lead) - 1
    agg = np.where(L.act & keep[root], cid[root], -1)
    return agg, int(lead.sum())


def coarsen(L, agg, n_c):
    """Galerkin operator on the aggregates: blocks summed in fine block order"""
    rows, cols = _rows(L), L.bnode
    I, J = agg[rows], np.where(L.act[cols], agg[cols], -1)
    m = (I >= 0) & (J >= 0)
    fb = np.flatnonzero(m)
    key = I[m] * n_c + J[m]
    order = np.argsort(key, kind="stable")
    key_s, fb_s = key[order], fb[order]
    head = np.ones(len(key_s), bool)
    head[1:] = key_s[1:] != key_s[:-1]
    starts = np.flatnonzero(head)
    lens = np.diff(np.append(starts, len(key_s)))
    vals = np.zeros((len(starts), 6))
    for p in range(int(lens.max()) if len(lens) else 0):          # sequential order inside every run
        mm = lens > p
        vals[mm] += L.bval[fb_s[starts[mm] + p]]
    C = Level()
    C.n = n_c
    ukey = key_s[starts]
    C.bnode = ukey % n_c
    C.brp = np.searchsorted(ukey // n_c, np.arange(n_c + 1)).astype(np.int64)
    C.bval = vals
    C.act = np.ones(n_c, bool)
    C.reg = L.reg
    owner = getattr(L, "owner", None)
    if owner is not None and n_c > REPLICATE_NODES:   # an aggregate lives on the rank of its members
        C.owner = np.zeros(n_c, dtype=owner.dtype)
        m = agg >= 0
        C.owner[agg[m]] = owner[m]
    return C


def diag_inverse(L):
    """(n,6) symmetric inverse of (diagonal block + reg I); zeros for inactive nodes / singular blocks"""
    rows = _rows(L)
    d = np.zeros((L.n, 6))
    m = L.bnode == rows
    d[rows[m]] = L.bval[m]
    xx, xy, xz, yy, yz, zz = (d[:, k] for k in range(6))
    xx = xx + L.reg; yy = yy + L.reg; zz = zz + L.reg
    c00 = yy * zz - yz * yz
    c01 = yz * xz - xy * zz
    c02 = xy * yz - yy * xz
    det = xx * c00 + xy * c01 + xz * c02
    ok = L.act & (det > 0) & np.isfinite(det)
    with np.errstate(all="ignore"):
        idet = np.where(ok, 1.0 / det, 0.0)
        inv = np.stack([c00, c01, c02, xx * zz - xz * xz, xz * xy - xx * yz, xx * yy - xy * xy], axis=1) * idet[:, None]
    inv[~ok] = 0.0
    return inv


def build_hierarchy(L0, verbose=False):
    levels = [L0]
    while True:
        L = levels[-1]
        L.dinv = diag_inverse(L)
        L.A = to_scipy(L)
        L.A_pc = to_scipy(L, PC_FP32)
        if len(levels) >= MAX_LEVELS or L.n <= MIN_NODES:
            break
        agg, n_c = aggregate(L)
        if n_c == 0 or n_c > MAX_RATIO * L.n:
            break
        L.agg, L.n_coarse = agg, n_c
        levels.append(coarsen(L, agg, n_c))
    if verbose:
        print("levels:", [(l.n, len(l.bnode)) for l in levels])
    return levels


def to_scipy(L, f32=False):
    """scalar CSR of the level operator restricted to active nodes, + reg I on active nodes.
    f32: block values rounded to FP32 first (what the V-cycle's sweeps stream on the GPU)."""
    v = L.bval.astype(np.float32).astype(np.float64) if f32 else L.bval
    data = np.stack([v[:, 0], v[:, 1], v[:, 2], v[:, 1], v[:, 3], v[:, 4], v[:, 2], v[:, 4], v[:, 5]], axis=1).reshape(-1, 3, 3)
    m = L.act[_rows(L)] & L.act[L.bnode]
    data = data * m[:, None, None]
    A = sp.bsr_matrix((data, L.bnode, L.brp), shape=(3 * L.n, 3 * L.n)).tocsr()
    return A + sp.diags(np.repeat(L.act, 3) * L.reg)


def apply_dinv(L, r):
    d = L.dinv
    r = r.reshape(-1, 3)
    out = np.empty_like(r)
    out[:, 0] = d[:, 0] * r[:, 0] + d[:, 1] * r[:, 1] + d[:, 2] * r[:, 2]
    out[:, 1] = d[:, 1] * r[:, 0] + d[:, 3] * r[:, 1] + d[:, 4] * r[:, 2]
    out[:, 2] = d[:, 2] * r[:, 0] + d[:, 4] * r[:, 1] + d[:, 5] * r[:, 2]
    return out.ravel()


def restrict(L, t):
    n_c = L.n_coarse
    out = np.zeros((n_c, 3))
    m = L.agg >= 0
    np.add.at(out, L.agg[m], t.reshape(-1, 3)[m])
    return out.ravel()


def prolong(L, ec):
    out = np.zeros((L.n, 3))
    m = L.agg >= 0
    out[m] = ec.reshape(-1, 3)[L.agg[m]]
    return out.ravel()


def vcycle(levels, l, r):
    L = levels[l]
    e = OMEGA * apply_dinv(L, r)
    if l == len(levels) - 1:
        for _ in range(COARSE_SWEEPS - 1):
            e = e + OMEGA * apply_dinv(L, r - L.A_pc @ e)
        return e
    t = r - L.A_pc @ e
    ec = vcycle(levels, l + 1, restrict(L, t))
    e = e + SCALE * prolong(L, ec)
    return e + OMEGA * apply_dinv(L, r - L.A_pc @ e)


def amg_pcg(K, free_mask, b, rtol=1e-10, maxit=5000, reg=1e-12, verbose=False, node_offsets=None):
    """PCG on A = K restricted to the free DOFs + reg I (rows/cols of known DOFs are zero, x = 0 there)
    with the V-cycle as M^-1.  Returns (x, iterations, levels).
    ``node_offsets`` (world+1,): the row partition of a multi-GPU solve -- rank r owns nodes
    [node_offsets[r], node_offsets[r+1]) and aggregates are formed inside a rank only."""
    L0 = level_from_csr(K, free_mask, reg)
    if L0 is None:
        raise ValueError("AMG needs node-complete Dirichlet sets")
    if node_offsets is not None:
        L0.owner = (np.searchsorted(np.asarray(node_offsets), np.arange(L0.n), side="right") - 1).astype(np.int32)
    levels = build_hierarchy(L0, verbose)
    A = levels[0].A
    fm = np.asarray(free_mask, bool)
    b = np.where(fm, b, 0.0)
    x = np.zeros_like(b)
    r = b.copy()
    bb = np.sqrt(b @ b)
    if bb == 0:
        return x, 0, levels
    z = vcycle(levels, 0, r)
    p = z.copy()
    rz = r @ z
    for it in range(1, maxit + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.sqrt(r @ r) <= rtol * bb:
            return x, it, levels
        z = vcycle(levels, 0, r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit, levels
