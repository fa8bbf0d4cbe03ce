"""Device-resident objects over the C-ABI: torch tensors are the buffers, ctypes the calls.

This is the layer the reference-compatible functions in ``fea_solver.py`` are written on.
Nothing here computes on the CPU; torch is used for allocation, streams and H2D/D2H copies.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

REGULARISATION = 1e-12      # src/fea_solver.py:125


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _hint(ctx, K):
    """Tell the library whether the CSR about to be passed has the node-block structure."""
    lib.myc_set_csr_hint(ctx.h, 1 if K.block3 else 0)


class Context:
    """One myc_ctx per device (scratch arenas, NCCL communicator)."""

    _by_device = {}

    def __init__(self, device_index: int):
        if not torch.cuda.is_available():
            raise RuntimeError("mycelium_fea_project_b200 needs a CUDA device (B200); there is no CPU path")
        h = C.c_void_p()
        rc = lib.myc_create(int(device_index), C.byref(h))
        if rc != 0:
            raise _lib.MyceliumFeaError(rc, (lib.myc_last_error(None) or b"").decode())
        self.h = h
        self.device = torch.device("cuda", int(device_index))
        self.rank, self.world = 0, 1
        self.node_offsets = None

    @classmethod
    def get(cls, device=None) -> "Context":
        idx = torch.cuda.current_device() if device is None else torch.device(device).index or 0
        if idx not in cls._by_device:
            cls._by_device[idx] = Context(idx)
        return cls._by_device[idx]

    @property
    def launches(self) -> int:
        return int(lib.myc_launch_count(self.h))

    def close(self):
        if self.h:
            lib.myc_destroy(self.h)
            self.h = None
        for k, v in list(Context._by_device.items()):
            if v is self:
                del Context._by_device[k]


@dataclass
class DeviceMesh:
    """A snapshot on the device: the reference's coords / elems / active (fea_solver.py:193-199)."""
    coords: torch.Tensor      # (n_nodes, 3) f64
    n1: torch.Tensor          # (n_elem,) i32
    n2: torch.Tensor          # (n_elem,) i32
    active: torch.Tensor      # (n_elem,) u8

    @property
    def n_nodes(self):
        return self.coords.shape[0]

    @property
    def n_elem(self):
        return self.n1.shape[0]

    @classmethod
    def from_host(cls, coords, n1, n2, active=None, device=None, pinned=False):
        dev = Context.get(device).device
        n_nodes = int(np.asarray(coords).shape[0])
        n1h = np.ascontiguousarray(n1)
        n2h = np.ascontiguousarray(n2)
        if n1h.size and (min(n1h.min(), n2h.min()) < 0 or max(n1h.max(), n2h.max()) >= n_nodes):
            raise IndexError("element end node outside [0, n_nodes)")     # numpy would raise too (:82-83)
        def up(a, dt):
            arr = np.ascontiguousarray(a, dtype=dt)
            t = torch.from_numpy(arr if arr.flags.writeable else arr.copy())
            if pinned:
                t = t.pin_memory()
            return t.to(dev, non_blocking=pinned)
        act = np.ones(len(n1h), dtype=np.uint8) if active is None else np.asarray(active).astype(np.uint8)
        return cls(up(np.asarray(coords, dtype=np.float64).reshape(-1, 3), np.float64),
                   up(n1h, np.int32), up(n2h, np.int32), up(act, np.uint8))


@dataclass
class DeviceCSR:
    """K (or a row block of it) in CSR on the device.  row_ptr local, col_idx global."""
    n_rows: int
    n_cols: int
    row_offset: int
    row_ptr: torch.Tensor     # (n_rows+1,) i32
    col_idx: torch.Tensor     # (nnz,) i32
    val: torch.Tensor         # (nnz,) f64
    block3: bool = False      # 3x3 node-block structure (always true for assembled K; verified otherwise)

    @property
    def nnz(self):
        return int(self.col_idx.shape[0])

    def to_scipy(self):
        from scipy.sparse import csr_matrix
        return csr_matrix((self.val.cpu().numpy(), self.col_idx.cpu().numpy(), self.row_ptr.cpu().numpy()),
                          shape=(self.n_rows, self.n_cols))

    @classmethod
    def from_scipy(cls, K, device=None):
        dev = Context.get(device).device
        K = K.tocsr()
        if not K.has_sorted_indices:
            K = K.sorted_indices()
        ctx = Context.get(device)
        out = cls(K.shape[0], K.shape[1], 0,
                  torch.from_numpy(K.indptr.astype(np.int32)).to(dev),
                  torch.from_numpy(K.indices.astype(np.int32)).to(dev),
                  torch.from_numpy(np.asarray(K.data, dtype=np.float64)).to(dev))
        flag = C.c_int(0)
        check(ctx.h, lib.myc_csr_is_block3(ctx.h, out.n_rows, _ptr(out.row_ptr), _ptr(out.col_idx), C.byref(flag),
                                           _stream()))
        out.block3 = bool(flag.value)
        return out


@dataclass
class DirichletSystem:
    ubc: torch.Tensor         # (n_cols,)  prescribed values scattered, 0 elsewhere
    rhs: torch.Tensor         # (n_rows,)  -K_fk u_k on free rows
    dinv: torch.Tensor        # (n_rows,)  1/(K_ii+reg) on free rows, 0 on known rows
    binv: torch.Tensor | None = None   # block-Jacobi inverses: (n_rows/3, 9) for "block3", symmetric-packed
                                       # (n_blocks, R(R+1)/2) for "block6" / "block12" (R = 6 / 12 rows per block)
    reg: float = REGULARISATION
    binv_kind: str | None = None       # which preconditioner ``binv`` belongs to
    precond: str = "jacobi"            # the preconditioner this system was prepared for ("amg" falls back to
                                       # "block6" when the multigrid hierarchy is not applicable, see apply_dirichlet)
    amg_levels: int = 0                # levels of the multigrid hierarchy (0: none)


# ---------------------------------------------------------------------------------------------
def bar_stiffness(ctx: Context, p1s: torch.Tensor, p2s: torch.Tensor, E, A, I):
    n = p1s.shape[0]
    K = torch.empty((n, 6, 6), dtype=torch.float64, device=ctx.device)
    L = torch.empty((n,), dtype=torch.float64, device=ctx.device)
    check(ctx.h, lib.myc_bar_stiffness_bulk(ctx.h, _ptr(p1s), _ptr(p2s), n, float(E), float(A), float(I),
                                            _ptr(K), _ptr(L), _stream()))
    return K, L


def assemble(ctx: Context, mesh: DeviceMesh, E, A, I, node_range=None) -> DeviceCSR:
    nb, ne = (0, mesh.n_nodes) if node_range is None else node_range
    n_rows = 3 * (ne - nb)
    row_ptr = torch.empty((n_rows + 1,), dtype=torch.int32, device=ctx.device)
    nnz = C.c_int64(0)
    check(ctx.h, lib.myc_assemble_symbolic(ctx.h, _ptr(mesh.n1), _ptr(mesh.n2), _ptr(mesh.active), mesh.n_elem,
                                           mesh.n_nodes, nb, ne, _ptr(row_ptr), C.byref(nnz), _stream()))
    col_idx = torch.empty((nnz.value,), dtype=torch.int32, device=ctx.device)
    val = torch.empty((nnz.value,), dtype=torch.float64, device=ctx.device)
    check(ctx.h, lib.myc_assemble_numeric(ctx.h, _ptr(mesh.coords), _ptr(mesh.n1), _ptr(mesh.n2), float(E), float(A),
                                          float(I), nnz.value, _ptr(row_ptr), _ptr(col_idx), _ptr(val), _stream()))
    return DeviceCSR(n_rows, 3 * mesh.n_nodes, 3 * nb, row_ptr, col_idx, val, block3=True)


def apply_dirichlet(ctx: Context, K: DeviceCSR, known_dofs: torch.Tensor, known_vals: torch.Tensor,
                    reg=REGULARISATION, block3=False, precond=None, reuse: DirichletSystem | None = None) -> DirichletSystem:
    """``precond`` ("jacobi", "block3", "block6", "block12", "amg") selects what is built next to the Jacobi
    diagonal: block inverses, or the aggregation-multigrid hierarchy (kept inside the context; it belongs to
    the LAST system prepared with precond="amg").  "amg" needs a node-block-structured, blockwise symmetric K
    and node-complete Dirichlet sets; otherwise the system is prepared for "block6" (``.precond`` tells).
    ``block3=True`` is the older spelling of precond="block3".
    ``reuse``: a system prepared earlier for the SAME K and the SAME set of known DOFs: only the prescribed
    values are new, so ubc / rhs are rewritten in place and the preconditioner is kept."""
    if reuse is not None:
        _hint(ctx, K)
        check(ctx.h, lib.myc_apply_dirichlet(ctx.h, K.n_rows, K.n_cols, K.row_offset, _ptr(K.row_ptr), _ptr(K.col_idx),
                                             _ptr(K.val), _ptr(known_dofs), _ptr(known_vals), known_dofs.shape[0],
                                             float(reuse.reg), _ptr(reuse.ubc), _ptr(reuse.rhs), _ptr(reuse.dinv), _stream()))
        return reuse
    if precond is None:
        precond = "block3" if block3 else "jacobi"
    if precond not in _lib.PRECONDITIONERS:
        raise ValueError(f"unknown preconditioner {precond!r}")
    _hint(ctx, K)
    ubc = torch.empty((K.n_cols,), dtype=torch.float64, device=ctx.device)
    rhs = torch.empty((K.n_rows,), dtype=torch.float64, device=ctx.device)
    dinv = torch.empty((K.n_rows,), dtype=torch.float64, device=ctx.device)
    check(ctx.h, lib.myc_apply_dirichlet(ctx.h, K.n_rows, K.n_cols, K.row_offset, _ptr(K.row_ptr), _ptr(K.col_idx),
                                         _ptr(K.val), _ptr(known_dofs), _ptr(known_vals), known_dofs.shape[0],
                                         float(reg), _ptr(ubc), _ptr(rhs), _ptr(dinv), _stream()))
    binv = None
    amg_levels = 0
    if precond == "amg":
        lv = C.c_int(0)
        check(ctx.h, lib.myc_amg_setup(ctx.h, K.n_rows, K.n_cols, K.row_offset, _ptr(K.row_ptr), _ptr(K.col_idx),
                                       _ptr(K.val), _ptr(dinv), float(reg), C.byref(lv), _stream()))
        amg_levels = int(lv.value)
        if amg_levels == 0:
            precond = "block6" if K.row_offset % 6 == 0 else "block3"
    if precond == "block3":
        binv = torch.empty((K.n_rows // 3, 9), dtype=torch.float64, device=ctx.device)
        check(ctx.h, lib.myc_block3_inverse(ctx.h, K.n_rows, K.row_offset, _ptr(K.row_ptr), _ptr(K.col_idx),
                                            _ptr(K.val), _ptr(dinv), float(reg), _ptr(binv), _stream()))
    elif precond in ("block6", "block12"):
        npb = 2 if precond == "block6" else 4
        R = 3 * npb
        n_blocks = (K.n_rows + R - 1) // R
        size = int(lib.myc_block_inverse_size(npb, K.n_rows))          # the layout belongs to the library
        binv = torch.empty((n_blocks, size // n_blocks if n_blocks else 0), dtype=torch.float64, device=ctx.device)
        check(ctx.h, lib.myc_block_inverse_packed(ctx.h, npb, K.n_rows, K.row_offset, _ptr(K.row_ptr), _ptr(K.col_idx),
                                                  _ptr(K.val), _ptr(dinv), float(reg), _ptr(binv), _stream()))
    return DirichletSystem(ubc, rhs, dinv, binv, float(reg), precond if binv is not None else None, precond,
                           amg_levels)


def spmv(ctx: Context, K: DeviceCSR, x: torch.Tensor, out: torch.Tensor | None = None):
    _hint(ctx, K)
    y = torch.empty((K.n_rows,), dtype=torch.float64, device=ctx.device) if out is None else out
    check(ctx.h, lib.myc_spmv(ctx.h, K.n_rows, _ptr(K.row_ptr), _ptr(K.col_idx), _ptr(K.val), _ptr(x), _ptr(y),
                              _stream()))
    return y


def pcg(ctx: Context, K: DeviceCSR, sys: DirichletSystem, x0: torch.Tensor | None = None, precond="jacobi",
        rtol=1e-10, atol=0.0, maxit=1_000_000, raise_on_maxit=True):
    """Returns (x, iterations, relres).  x is the solution on free rows (0 on known rows)."""
    x = torch.zeros((K.n_rows,), dtype=torch.float64, device=ctx.device) if x0 is None else x0
    if precond == "amg" and sys.precond != "amg":
        precond = sys.precond                          # hierarchy not applicable: the system was prepared for a fallback
    pc = _lib.PRECONDITIONERS[precond]
    if pc == _lib.MYC_PC_AMG:
        if sys.precond != "amg":
            raise ValueError("amg preconditioner needs apply_dirichlet(..., precond='amg')")
    elif pc != _lib.MYC_PC_JACOBI and (sys.binv is None or sys.binv_kind != precond):
        raise ValueError(f"{precond} preconditioner needs apply_dirichlet(..., precond={precond!r})")
    iters, relres = C.c_int64(0), C.c_double(0.0)
    _hint(ctx, K)
    rc = lib.myc_pcg_solve(ctx.h, K.n_rows, K.n_cols, K.row_offset, _ptr(K.row_ptr), _ptr(K.col_idx), _ptr(K.val),
                           _ptr(sys.rhs), _ptr(sys.dinv), _ptr(sys.binv), pc, float(sys.reg), float(rtol),
                           float(atol), int(maxit), _ptr(x), C.byref(iters), C.byref(relres), _stream())
    check(ctx.h, rc, allow_not_converged=not raise_on_maxit)
    return x, int(iters.value), float(relres.value)


def amg_levels(ctx: Context, detail=False):
    """[(nodes, blocks)] per level of the context's current multigrid hierarchy (what THIS rank holds), and its
    setup time in ms.  ``detail``: dicts with n, nb, node_off, n_global, replicated (0 partitioned over the ranks /
    single GPU, 1 first replicated level, 2 replicated) and agg_shift instead of the tuples."""
    out = (C.c_int64 * 8)()
    check(ctx.h, lib.myc_amg_level_info(ctx.h, 0, out, None, _stream()))
    n_levels, setup_ms = int(out[2]), out[3] / 1e3
    levels = []
    for l in range(n_levels):
        check(ctx.h, lib.myc_amg_level_info(ctx.h, l, out, None, _stream()))
        if detail:
            levels.append({"n": int(out[0]), "nb": int(out[1]), "node_off": int(out[4]), "n_global": int(out[5]),
                           "replicated": int(out[6]), "agg_shift": int(out[7])})
        else:
            levels.append((int(out[0]), int(out[1])))
    return levels, setup_ms


def amg_aggregates(ctx: Context, level: int) -> torch.Tensor:
    """node -> aggregate map of ``level`` (int32, -1 = not represented on the next level) for the nodes this rank
    holds; values index the next level from the first node this rank holds there (see amg_levels(detail=True))."""
    out = (C.c_int64 * 8)()
    check(ctx.h, lib.myc_amg_level_info(ctx.h, level, out, None, _stream()))
    agg = torch.empty((int(out[0]),), dtype=torch.int32, device=ctx.device)
    check(ctx.h, lib.myc_amg_level_info(ctx.h, level, out, _ptr(agg), _stream()))
    return agg


def true_residual(ctx: Context, K: DeviceCSR, sys: DirichletSystem, x: torch.Tensor) -> float:
    out = C.c_double(0.0)
    _hint(ctx, K)
    check(ctx.h, lib.myc_true_residual(ctx.h, K.n_rows, K.n_cols, K.row_offset, _ptr(K.row_ptr), _ptr(K.col_idx),
                                       _ptr(K.val), _ptr(sys.rhs), _ptr(sys.dinv), float(sys.reg), _ptr(x),
                                       C.byref(out), _stream()))
    return float(out.value)


def merge_solution(ctx: Context, K: DeviceCSR, sys: DirichletSystem, x: torch.Tensor, out: torch.Tensor | None = None):
    """U (global length): x on free rows, prescribed values on known rows (fea_solver.py:131-133)."""
    U = torch.zeros((K.n_cols,), dtype=torch.float64, device=ctx.device) if out is None else out
    check(ctx.h, lib.myc_merge_solution(ctx.h, K.n_rows, K.row_offset, _ptr(x), _ptr(sys.dinv), _ptr(sys.ubc),
                                        _ptr(U), _stream()))
    return U


def gather_sum(ctx: Context, v: torch.Tensor, idx: torch.Tensor) -> float:
    out = C.c_double(0.0)
    check(ctx.h, lib.myc_gather_sum(ctx.h, _ptr(v), _ptr(idx), idx.shape[0], C.byref(out), _stream()))
    return float(out.value)


def strain_update(ctx: Context, mesh: DeviceMesh, U: torch.Tensor, E, max_strain):
    """In-place failure update of mesh.active; returns (stress tensor, n_active)."""
    stress = torch.empty((mesh.n_elem,), dtype=torch.float64, device=ctx.device)
    n_act = C.c_int64(0)
    check(ctx.h, lib.myc_strain_update(ctx.h, _ptr(mesh.coords), _ptr(mesh.n1), _ptr(mesh.n2), mesh.n_elem, _ptr(U),
                                       float(E), float(max_strain), _ptr(mesh.active), _ptr(stress),
                                       C.byref(n_act), _stream()))
    return stress, int(n_act.value)


def reduce_csr(ctx: Context, K: DeviceCSR, sys: DirichletSystem):
    """Explicit K[free][:,free] (structure-parity aid, fea_solver.py:118).  Single GPU."""
    free_index = torch.empty((K.n_rows + 1,), dtype=torch.int32, device=ctx.device)
    rrp = torch.empty((K.n_rows + 1,), dtype=torch.int32, device=ctx.device)
    n_free, nnz = C.c_int64(0), C.c_int64(0)
    check(ctx.h, lib.myc_reduce_csr(ctx.h, K.n_rows, _ptr(K.row_ptr), _ptr(K.col_idx), _ptr(K.val), _ptr(sys.dinv),
                                    _ptr(free_index), _ptr(rrp), None, None, C.byref(n_free), C.byref(nnz),
                                    _stream()))
    rci = torch.empty((nnz.value,), dtype=torch.int32, device=ctx.device)
    rv = torch.empty((nnz.value,), dtype=torch.float64, device=ctx.device)
    nnz2 = C.c_int64(0)
    check(ctx.h, lib.myc_reduce_csr(ctx.h, K.n_rows, _ptr(K.row_ptr), _ptr(K.col_idx), _ptr(K.val), _ptr(sys.dinv),
                                    _ptr(free_index), _ptr(rrp), _ptr(rci), _ptr(rv), C.byref(n_free),
                                    C.byref(nnz2), _stream()))
    nf = int(n_free.value)
    return DeviceCSR(nf, nf, 0, rrp[:nf + 1], rci, rv, block3=False)
