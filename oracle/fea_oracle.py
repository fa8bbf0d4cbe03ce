"""numpy/scipy restatement of the reference FEA path -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines (under /root/reference/) it restates.
The arithmetic follows the reference operation for operation (same numpy ufuncs in
the same order, same scipy calls) so that results are bit-identical to the
reference on this container; only the Python-level 36-append triple loop of
``assemble_global_stiffness`` (src/fea_solver.py:93-103) is replaced by index
arithmetic that emits the very same COO triplets in the very same order.

See oracle/__init__.py for what may import this file and how it is pinned.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix, identity
from scipy.sparse.linalg import spsolve

# ---------------------------------------------------------------------------
# Constants -- src/fea_solver.py:14-28 (same expressions, so the same doubles).
# ---------------------------------------------------------------------------
E_mod = 2500
d = 0.0002
t = 0.000001
A = 3.14 * ((d / 2) ** 2 - (d / 2 - t) ** 2)      # 3.14, not pi (fea_solver.py:17)
I = A * 0.001
N_STEPS = 40
DISPLACEMENT_MAX = 0.02
MAX_STRAIN = 0.018
MAX_STRESS = E_mod * MAX_STRAIN
GRIP_LENGTH = 1.5
REGULARISATION = 1e-12                               # fea_solver.py:125


# ---------------------------------------------------------------------------
# a3: element stiffness -- src/fea_solver.py:30-68
# ---------------------------------------------------------------------------
def bar_stiffness_bulk(p1s, p2s, E=E_mod, A=A, I=I):
    """K_e (N,6,6) and L (N,) for 2-node bars; restates fea_solver.py:30-68.

    S = (n n^T) k_ax + (I3 - n n^T) k_b laid out as [[S,-S],[-S,S]], with each
    product rounded before the add, L clamped below at 1e-12 (:41).
    """
    p1s = np.asarray(p1s, dtype=np.float64)
    p2s = np.asarray(p2s, dtype=np.float64)
    vec = p2s - p1s                                        # :37
    L = np.linalg.norm(vec, axis=1)                        # :38
    L_safe = np.where(L < 1e-12, 1e-12, L)                 # :41
    n = vec / L_safe[:, None]                              # :42
    k_ax = (E * A) / L_safe                                # :45
    col = n[:, :, None]                                    # :46
    nnT = col @ col.transpose(0, 2, 1)                     # :47
    N = len(L)
    K_ax = np.zeros((N, 6, 6))
    K_ax[:, 0:3, 0:3] = nnT                                # :50-53
    K_ax[:, 0:3, 3:6] = -nnT
    K_ax[:, 3:6, 0:3] = -nnT
    K_ax[:, 3:6, 3:6] = nnT
    K_ax *= k_ax[:, None, None]                            # :54
    perp = np.eye(3) - n[:, None, :] * n[:, :, None]       # :57
    k_b = 12 * E * I / (L_safe ** 3)                       # :58
    K_b = np.zeros((N, 6, 6))
    K_b[:, 0:3, 0:3] = perp                                # :62-65
    K_b[:, 3:6, 0:3] = -perp
    K_b[:, 0:3, 3:6] = -perp
    K_b[:, 3:6, 3:6] = perp
    K_b *= k_b[:, None, None]                              # :66
    return K_ax + K_b, L                                   # :68


# ---------------------------------------------------------------------------
# a4: assembly -- src/fea_solver.py:74-106
# ---------------------------------------------------------------------------
def coo_triplets(n1, n2, K_e_all):
    """The COO stream of fea_solver.py:93-103: element-major, i-major, j-minor."""
    n1 = np.asarray(n1, dtype=np.int64)
    n2 = np.asarray(n2, dtype=np.int64)
    off = np.arange(3, dtype=np.int64)
    dof = np.concatenate([3 * n1[:, None] + off, 3 * n2[:, None] + off], axis=1)  # :96
    rows = np.repeat(dof, 6, axis=1).ravel()               # dof[i] for each of 6 j  (:101)
    cols = np.tile(dof, (1, 6)).ravel()                    # dof[j]                  (:102)
    vals = np.ascontiguousarray(K_e_all).reshape(-1)       # Ke[i, j]                (:103)
    return rows, cols, vals


def assemble_global_stiffness(coords, n1, n2, active, E=E_mod, A=A, I=I):
    """Global K as scipy CSR; restates fea_solver.py:74-106.

    ``n1``/``n2`` are the element end-node columns (``elems["n1"]``, ``elems["n2"]``).
    scipy's COO constructor sums duplicates, sorts columns and keeps explicit zeros.
    """
    coords = np.asarray(coords, dtype=np.float64)
    n_dof = 3 * coords.shape[0]                            # :75-76
    eidx = np.where(np.asarray(active))[0]                 # :79
    a1 = np.asarray(n1)[eidx]
    a2 = np.asarray(n2)[eidx]
    K_e_all, _ = bar_stiffness_bulk(coords[a1], coords[a2], E, A, I)   # :82-86
    rows, cols, vals = coo_triplets(a1, a2, K_e_all)
    return csr_matrix((vals, (rows, cols)), shape=(n_dof, n_dof))      # :105


def assemble_global_stiffness_loop(coords, n1, n2, active, E=E_mod, A=A, I=I):
    """Same as above but with the reference's literal append loop (small meshes only)."""
    coords = np.asarray(coords, dtype=np.float64)
    n_dof = 3 * coords.shape[0]
    eidx = np.where(np.asarray(active))[0]
    a1 = np.asarray(n1)[eidx]
    a2 = np.asarray(n2)[eidx]
    K_e_all, _ = bar_stiffness_bulk(coords[a1], coords[a2], E, A, I)
    rows, cols, vals = [], [], []
    for k in range(len(eidx)):
        dof = np.r_[3 * int(a1[k]):3 * int(a1[k]) + 3, 3 * int(a2[k]):3 * int(a2[k]) + 3]
        Ke = K_e_all[k]
        for i in range(6):
            for j in range(6):
                rows.append(dof[i]); cols.append(dof[j]); vals.append(Ke[i, j])
    return csr_matrix((vals, (rows, cols)), shape=(n_dof, n_dof))


# ---------------------------------------------------------------------------
# a5: grips and Dirichlet sets -- src/fea_solver.py:205-210, 223-245
# ---------------------------------------------------------------------------
def grip_nodes(coords, tol=GRIP_LENGTH, axis=1):
    """(hi_nodes, lo_nodes): nodes within ``tol`` of the max / min coordinate along
    ``axis``; axis=1 is the reference's top/bottom grips (fea_solver.py:205-210)."""
    c = np.asarray(coords)[:, axis]
    c_min, c_max = c.min(), c.max()
    ids = np.arange(len(c))
    hi = ids[np.abs(c - c_max) < tol].astype(int)
    lo = ids[np.abs(c - c_min) < tol].astype(int)
    return hi, lo


def build_bc(hi_nodes, lo_nodes, d_hi, d_lo, comp=1):
    """known_dofs (dict insertion order) and known_vals; restates fea_solver.py:223-245.

    ``comp`` is the prescribed component (1 = y, the reference's only load case); the
    other two components of every grip node are clamped to 0.  A node in both sets keeps
    the position of its first insertion and the value of the last (dict semantics).
    """
    disp = {}
    for nodes, val in ((hi_nodes, d_hi), (lo_nodes, d_lo)):
        for n in nodes:
            n = int(n)
            disp.update({3 * n + c: (val if c == comp else 0.0) for c in range(3)})
    known_dofs = np.array(list(disp.keys()), dtype=np.int64)
    known_vals = np.array([disp[k] for k in known_dofs], dtype=np.float64)
    return known_dofs, known_vals


def ramp_displacements(step, n_steps=N_STEPS, disp_max=DISPLACEMENT_MAX):
    """(dy_top, dy_bot) of ramp step ``step``; fea_solver.py:217-219."""
    f = step / (n_steps - 1)
    return +disp_max * f, -disp_max * f


# ---------------------------------------------------------------------------
# a6/a7: reduction + direct solve -- src/fea_solver.py:112-135
# ---------------------------------------------------------------------------
def reduce_system(K, known_dofs, known_vals):
    """(free_dofs, K_ff + 1e-12 I, F_f); restates fea_solver.py:113-125."""
    n_dof = K.shape[0]
    free = np.setdiff1d(np.arange(n_dof), known_dofs)      # :115
    K_ff = K[free][:, free].tocsr()                        # :118
    K_fk = K[free][:, known_dofs]                          # :119
    F = np.zeros(n_dof)
    F_f = F[free] - K_fk @ known_vals                      # :121-122
    K_ff = K_ff + REGULARISATION * identity(K_ff.shape[0], format="csr")   # :125
    return free, K_ff, F_f


def solve_system(K, known_dofs, known_vals):
    """U (n_dof,) by sparse direct solve; restates fea_solver.py:112-135."""
    n_dof = K.shape[0]
    free, K_ff, F_f = reduce_system(K, known_dofs, known_vals)
    U_f = spsolve(K_ff, F_f)                               # :128
    U = np.zeros(n_dof)
    U[free] = U_f                                          # :131-133
    U[known_dofs] = known_vals
    return U


# ---------------------------------------------------------------------------
# a8 + strain/failure -- src/fea_solver.py:257-284
# ---------------------------------------------------------------------------
def reactions(K, U, top_nodes):
    """Sum of y reactions on the top grip; fea_solver.py:257,263-264."""
    F = K @ U
    return F[[3 * int(n) + 1 for n in top_nodes]].sum()


def strain_stress_update(coords, n1, n2, U, active, E=E_mod, max_strain=MAX_STRAIN):
    """Per-element axial stress and in-place failure update; fea_solver.py:269-284.

    Element by element like the reference (np.linalg.norm / np.dot on 3-vectors,
    L NOT clamped -- :276-277).
    """
    stress = np.zeros(len(n1))
    for i in range(len(n1)):
        if not active[i]:
            continue
        a, b = int(n1[i]), int(n2[i])
        vec = coords[b] - coords[a]
        L = np.linalg.norm(vec)
        n = vec / L
        strain = np.dot(n, U[3 * b:3 * b + 3] - U[3 * a:3 * a + 3]) / L
        stress[i] = E * strain
        if abs(strain) > max_strain:
            active[i] = False
    return stress


# ---------------------------------------------------------------------------
# Driver -- src/fea_solver.py:186-335 (no plotting variant)
# ---------------------------------------------------------------------------
@dataclass
class RampResult:
    stress: list = field(default_factory=list)
    active: list = field(default_factory=list)
    disp: list = field(default_factory=list)
    force_disp: list = field(default_factory=list)


def fea_ramp(coords, n1, n2, tol=GRIP_LENGTH, n_steps=N_STEPS, disp_max=DISPLACEMENT_MAX,
             max_strain=MAX_STRAIN, E=E_mod, A=A, I=I, solve=solve_system):
    """The displacement-ramp loop of fea_solver.py:213-295 on in-memory arrays."""
    coords = np.asarray(coords, dtype=np.float64)
    n_elems = len(n1)
    active = np.ones(n_elems, dtype=bool)
    top, bot = grip_nodes(coords, tol, axis=1)
    out = RampResult()
    for step in range(n_steps):
        dy_top, dy_bot = ramp_displacements(step, n_steps, disp_max)
        K = assemble_global_stiffness(coords, n1, n2, active, E, A, I)
        known_dofs, known_vals = build_bc(top, bot, dy_top, dy_bot, comp=1)
        try:
            U = solve(K, known_dofs, known_vals)
        except np.linalg.LinAlgError:
            break
        out.force_disp.append([dy_top - dy_bot, reactions(K, U, top)])
        stress = strain_stress_update(coords, n1, n2, U, active, E, max_strain)
        out.stress.append(stress)
        out.active.append(active.copy())
        out.disp.append(U.copy())
        if active.sum() == 0:
            break
    return out


def write_results(fea_dir, res: RampResult, n_elems, runtime_s=None):
    """The four CSVs of fea_solver.py:298-316 (+ runtime.txt :331-333)."""
    os.makedirs(fea_dir, exist_ok=True)
    cols = [f"elem_{i}" for i in range(n_elems)]
    s = pd.DataFrame(res.stress, columns=cols)
    s["step"] = np.arange(1, len(res.stress) + 1)
    s.to_csv(os.path.join(fea_dir, "stress_record.csv"), index=False)
    a = pd.DataFrame(res.active, columns=cols)
    a["step"] = np.arange(1, len(res.active) + 1)
    a.to_csv(os.path.join(fea_dir, "active_elements.csv"), index=False)
    dd = pd.DataFrame(res.disp, columns=np.arange(len(res.disp[0])))
    dd["step"] = np.arange(1, len(res.disp) + 1)
    dd.to_csv(os.path.join(fea_dir, "node_displacements.csv"), index=False)
    fd = pd.DataFrame(res.force_disp, columns=["total_displacement", "total_force"])
    fd.to_csv(os.path.join(fea_dir, "force_displacement.csv"), index=False)
    if runtime_s is not None:
        with open(os.path.join(fea_dir, "runtime.txt"), "w") as f:
            f.write(f"Total FEA runtime: {runtime_s:.6f} seconds\n")


def fea_solver(results_dir, tol=GRIP_LENGTH, **kw):
    """CSV-in / CSV-out entry point, fea_solver.py:186."""
    t0 = time.time()
    nodes = pd.read_csv(os.path.join(results_dir, "nodes.csv"))
    elems = pd.read_csv(os.path.join(results_dir, "elements.csv"))
    coords = nodes[["x", "y", "z"]].values
    res = fea_ramp(coords, elems["n1"].values, elems["n2"].values, tol=tol, **kw)
    write_results(os.path.join(results_dir, "fea_results"), res, len(elems), time.time() - t0)
    return res


# ---------------------------------------------------------------------------
# Iterative baseline: the algorithm PETSc's KSPCG + PCJACOBI runs on the reduced
# system (src/fea_petsc.cpp:323-341 with -pc_type jacobi; the solver x PC menu of
# src/fea_petsc_solverAndPC.cpp:330-331).  PETSc itself is absent from this image,
# so this is a restatement of the published algorithm (Hestenes-Stiefel PCG, left
# preconditioning, x0 = 0), not a run of PETSc 3.24.1.  Convergence is tested on
# the unpreconditioned residual ||r||2 <= rtol*||b||2 (KSP_NORM_UNPRECONDITIONED).
# ---------------------------------------------------------------------------
def jacobi_pcg(Aop, b, rtol=1e-10, maxit=200000, block3=None):
    """Returns (x, iterations, ||r||/||b||).  ``Aop`` is a scipy CSR (with the 1e-12
    shift already added)."""
    n = len(b)
    dinv = 1.0 / Aop.diagonal()
    x = np.zeros(n)
    r = b.copy()
    bnorm = np.linalg.norm(b)
    if bnorm == 0.0:
        return x, 0, 0.0
    z = dinv * r
    p = z.copy()
    rz = r @ z
    it = 0
    rn = bnorm
    while it < maxit:
        Ap = Aop @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        it += 1
        rn = np.linalg.norm(r)
        if rn <= rtol * bnorm:
            break
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it, rn / bnorm


def solve_system_pcg(K, known_dofs, known_vals, rtol=1e-10, maxit=200000):
    """solve_system with the direct solve swapped for Jacobi-PCG (CPU iterative baseline)."""
    n_dof = K.shape[0]
    free, K_ff, F_f = reduce_system(K, known_dofs, known_vals)
    U_f, it, rel = jacobi_pcg(K_ff.tocsr(), F_f, rtol, maxit)
    U = np.zeros(n_dof)
    U[free] = U_f
    U[known_dofs] = known_vals
    return U, it, rel

# ---------------------------------------------------------------------------
# Checker for the node-group block-Jacobi preconditioners of the CUDA path (MYC_PC_BLOCK3 / 6 / 12):
# PETSc's PCBJACOBI idea (src/fea_petsc_parallel.cpp:339) with fixed aligned blocks of
# ``rows_per_block`` consecutive DOFs of the FULL index space, restricted to the free DOFs, inverted
# exactly.  Not in the reference's scipy path; used only to predict iteration counts and block inverses.
# ---------------------------------------------------------------------------
def aligned_block_inverses(K_ff, free, rows_per_block):
    """Dense inverses of the diagonal blocks of K_ff grouped by ``free // rows_per_block``.
    Returns (labels (n_free,), {label: (reduced row indices, inverse)})."""
    labels = np.asarray(free) // rows_per_block
    A = K_ff.tocsr()
    out = {}
    starts = np.flatnonzero(np.r_[True, labels[1:] != labels[:-1]])
    ends = np.r_[starts[1:], len(labels)]
    for s, e in zip(starts, ends):
        idx = np.arange(s, e)
        out[int(labels[s])] = (idx, np.linalg.inv(A[s:e, s:e].toarray()))
    return labels, out


def block_jacobi_pcg(K_ff, F_f, free, rows_per_block, rtol=1e-10, maxit=200000):
    """Hestenes-Stiefel PCG on the reduced system with the aligned-block Jacobi preconditioner.
    Returns (x, iterations, ||r||/||b||)."""
    A = K_ff.tocsr()
    _, blocks = aligned_block_inverses(A, free, rows_per_block)
    # assemble the block-diagonal inverse once as a sparse matrix
    rows, cols, vals = [], [], []
    for idx, inv in blocks.values():
        r, c = np.meshgrid(idx, idx, indexing="ij")
        rows.append(r.ravel()); cols.append(c.ravel()); vals.append(inv.ravel())
    Minv = csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=A.shape)
    b = np.asarray(F_f, dtype=float)
    x = np.zeros(len(b))
    bnorm = np.linalg.norm(b)
    if bnorm == 0.0:
        return x, 0, 0.0
    r = b.copy()
    z = Minv @ r
    p = z.copy()
    rz = r @ z
    it, rn = 0, bnorm
    while it < maxit:
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        it += 1
        rn = np.linalg.norm(r)
        if rn <= rtol * bnorm:
            break
        z = Minv @ r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it, rn / bnorm
