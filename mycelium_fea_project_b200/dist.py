"""Row-partitioned multi-GPU solve: one process per GPU, NCCL over NVLink.

Partition (the PETSc MPIAIJ layout of src/fea_petsc_parallel.cpp:236, at node granularity):
rank r owns the contiguous node range [offsets[r], offsets[r+1]) and therefore DOF rows
[3*offsets[r], 3*offsets[r+1]).  Each rank assembles ONLY its rows, from the elements incident
to its nodes (elements crossing a cut are evaluated on both sides; K_e is cheap), so assembly
needs no communication -- and none of the reference's P-fold over-assembly
(src/fea_petsc_parallel.cpp:242-265).

Vectors gathered through column indices are kept at global length on every rank, so a halo
refresh is a send/recv of contiguous DOF ranges between the peers' buffers at identical
offsets (csrc/dist.cu).  The plan below (which ranges) is plain numpy and is exercised on CPU
with gloo (tests/test_dist_cpu.py); the exchange inside the solver runs in the C library on
NCCL.  Meshes whose node numbering is not spatially local should be renumbered first
(``locality_order``), otherwise the ranges degenerate to whole partitions.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np


# ---------------------------------------------------------------------------------------------
# plan (pure numpy -- no CUDA, no torch)
# ---------------------------------------------------------------------------------------------
def partition_nodes(n_nodes: int, world: int, align: int = 1) -> np.ndarray:
    """Balanced contiguous node ranges: offsets[world+1] (first n_nodes % world ranks get one more,
    PETSC_DECIDE's rule applied to nodes instead of rows).  ``align`` > 1 puts every interior cut on a
    multiple of ``align`` nodes (the aligned block-Jacobi groups must not straddle ranks); the balance
    rule is then applied to groups of ``align`` nodes and the last rank takes the ragged tail."""
    n_nodes, world, align = int(n_nodes), int(world), int(align)
    if align <= 1:
        base, rem = divmod(n_nodes, world)
        sizes = np.full(world, base, dtype=np.int64)
        sizes[:rem] += 1
        return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    groups = -(-n_nodes // align)
    off = np.minimum(partition_nodes(groups, world) * align, n_nodes)
    off[-1] = n_nodes
    return off.astype(np.int64)


def owner_of(nodes, offsets):
    return np.searchsorted(offsets, nodes, side="right") - 1


def halo_ranges(n1, n2, active, offsets, rank):
    """For ``rank``: per peer q the half-open node range [lo, hi) of q's nodes that appear as
    the far end of an active element with a near end owned by ``rank`` (lo == hi: nothing)."""
    world = len(offsets) - 1
    n1 = np.asarray(n1, dtype=np.int64)
    n2 = np.asarray(n2, dtype=np.int64)
    if active is not None:
        m = np.asarray(active).astype(bool)
        n1, n2 = n1[m], n2[m]
    lo = np.zeros(world, dtype=np.int64)
    hi = np.zeros(world, dtype=np.int64)
    b, e = offsets[rank], offsets[rank + 1]
    far = np.concatenate([n2[(n1 >= b) & (n1 < e)], n1[(n2 >= b) & (n2 < e)]])
    far = far[(far < b) | (far >= e)]
    if far.size:
        own = owner_of(far, offsets)
        for q in np.unique(own):
            f = far[own == q]
            lo[q], hi[q] = f.min(), f.max() + 1
    return lo, hi


def locality_order(coords, axis=1):
    """Permutation that numbers nodes along ``axis`` (ties by the other in-plane axis) so that
    contiguous ranges are spatial strips.  perm[new] = old."""
    c = np.asarray(coords)
    other = 0 if axis == 1 else 1
    return np.lexsort((c[:, other], c[:, axis]))


@dataclass
class HaloPlan:
    rank: int
    world: int
    offsets: np.ndarray      # (world+1,) node offsets
    need_lo: np.ndarray      # (world,) node ranges this rank receives from q
    need_hi: np.ndarray
    give_lo: np.ndarray      # (world,) node ranges this rank sends to q
    give_hi: np.ndarray

    @property
    def node_begin(self):
        return int(self.offsets[self.rank])

    @property
    def node_end(self):
        return int(self.offsets[self.rank + 1])


def make_plan(n1, n2, active, n_nodes, rank, world, all_gather=None, align=1) -> HaloPlan:
    """Build this rank's plan.  ``all_gather(array) -> list of arrays`` exchanges the need
    table between ranks (torch.distributed in production); if None, every rank's needs are
    computed locally (all ranks hold the whole mesh, so this is equivalent, just O(world) more
    host work)."""
    offsets = partition_nodes(n_nodes, world, align)
    lo, hi = halo_ranges(n1, n2, active, offsets, rank)
    if all_gather is not None:
        table = all_gather(np.stack([lo, hi]))
    else:
        table = [np.stack(halo_ranges(n1, n2, active, offsets, q)) for q in range(world)]
    give_lo = np.array([table[q][0][rank] for q in range(world)], dtype=np.int64)
    give_hi = np.array([table[q][1][rank] for q in range(world)], dtype=np.int64)
    return HaloPlan(rank, world, offsets, lo, hi, give_lo, give_hi)


def exchange_halo_torch(x_global, plan: HaloPlan, group=None):
    """Reference halo refresh with torch.distributed point-to-point ops (works on gloo and
    nccl).  Same semantics as myc_halo_exchange; used by the CPU tests."""
    import torch.distributed as dist
    ops = []
    for q in range(plan.world):
        if q == plan.rank:
            continue
        if plan.give_hi[q] > plan.give_lo[q]:
            ops.append(dist.P2POp(dist.isend, x_global[3 * plan.give_lo[q]:3 * plan.give_hi[q]], q, group))
        if plan.need_hi[q] > plan.need_lo[q]:
            ops.append(dist.P2POp(dist.irecv, x_global[3 * plan.need_lo[q]:3 * plan.need_hi[q]], q, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return x_global


# ---------------------------------------------------------------------------------------------
# production path: C library + NCCL
# ---------------------------------------------------------------------------------------------
def local_mesh_slice(n1, n2, node_begin, node_end):
    """What rank [node_begin, node_end) needs of the mesh: the indices of the elements incident to an owned node
    (in element order, so duplicate elements are summed in the reference's order) and the node range
    [lo, hi) their end nodes span (owned nodes + the halo the cut elements reach)."""
    n1 = np.asarray(n1)
    n2 = np.asarray(n2)
    m = ((n1 >= node_begin) & (n1 < node_end)) | ((n2 >= node_begin) & (n2 < node_end))
    idx = np.flatnonzero(m)
    lo, hi = int(node_begin), int(node_end)
    if idx.size:
        lo = min(lo, int(n1[idx].min()), int(n2[idx].min()))
        hi = max(hi, int(n1[idx].max()) + 1, int(n2[idx].max()) + 1)
    return idx, lo, hi


def shutdown(ctx=None):
    """Orderly end of a multi-GPU job (collective; call before ``dist.destroy_process_group()``): every rank unmaps
    its peers' solver buffers, all ranks meet, then every rank destroys its context (which frees the buffers it had
    exported) -- the order CUDA IPC asks for."""
    import torch
    import torch.distributed as dist
    from . import device as dv
    from ._lib import lib
    ctx = dv.Context.get() if ctx is None else ctx
    torch.cuda.synchronize()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
        lib.myc_dist_release_peers(ctx.h)
        dist.barrier()
    ctx.close()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


class DistributedSolver:
    """Collective driver of one load case on a row-partitioned mesh.  The constructor sees the whole mesh
    description on the host (it is what a snapshot reader hands over), but each rank uploads and keeps only its
    share: the elements incident to its nodes and the coordinates of its nodes + halo; matrices and solver
    vectors are partitioned (src/fea_petsc_parallel.cpp:236 -- without its every-rank-assembles-everything loop,
    :242-265)."""

    def __init__(self, mesh_host, active=None, device=None):
        import torch
        import torch.distributed as dist
        from . import device as dv
        from ._lib import lib, check, nccl_library_path
        coords, n1, n2 = mesh_host
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.ctx = dv.Context.get(device)
        self.n_nodes = int(len(coords))

        def all_gather(arr):
            out = [None] * self.world
            dist.all_gather_object(out, arr)
            return out

        # cuts on even nodes, so that the aligned 6x6 Jacobi blocks stay rank-local ("block6" on N > 1 GPUs)
        self.block6 = True
        self.plan = make_plan(n1, n2, active, self.n_nodes, self.rank, self.world, all_gather, align=2)
        # this rank's share of the mesh.  Node ids stay global (the assembler indexes coords with them), so the
        # coordinate array keeps its global length on the device, but only [coord_lo, coord_hi) is ever written or read.
        self.elem_index, self.coord_lo, self.coord_hi = local_mesh_slice(n1, n2, self.plan.node_begin, self.plan.node_end)
        n1h, n2h = np.asarray(n1), np.asarray(n2)
        if n1h.size and (min(n1h.min(), n2h.min()) < 0 or max(n1h.max(), n2h.max()) >= self.n_nodes):
            raise IndexError("element end node outside [0, n_nodes)")
        dev = self.ctx.device
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        coords_d = torch.empty((self.n_nodes, 3), dtype=torch.float64, device=dev)
        coords_d[self.coord_lo:self.coord_hi] = up(np.asarray(coords, dtype=np.float64).reshape(-1, 3)[self.coord_lo:self.coord_hi], np.float64)
        act = np.ones(len(self.elem_index), np.uint8) if active is None else np.asarray(active)[self.elem_index].astype(np.uint8)
        self.mesh = dv.DeviceMesh(coords_d, up(n1h[self.elem_index], np.int32), up(n2h[self.elem_index], np.int32), up(act, np.uint8))
        if self.world > 1 and self.ctx.world == 1:
            path = nccl_library_path().encode()
            uid = np.zeros(128, dtype=np.uint8)
            if self.rank == 0:
                rc = lib.myc_dist_unique_id(path, uid.ctypes.data_as(C.c_void_p))
                if rc != 0:
                    raise RuntimeError("myc_dist_unique_id failed (NCCL not loadable)")
            t = torch.from_numpy(uid).to(self.ctx.device)
            dist.broadcast(t, 0)
            uid = t.cpu().numpy()
            check(self.ctx.h, lib.myc_dist_init(self.ctx.h, path, uid.ctypes.data_as(C.c_void_p), self.rank,
                                                self.world))
            self.ctx.rank, self.ctx.world = self.rank, self.world
        self._install_plan()
        # NVLink peer buffers now, not inside the first solve: opening 7 IPC handles (peer access is
        # enabled lazily per peer) takes seconds on an 8-GPU board
        self._ensure_peer(3 * self.n_nodes)

    def _install_plan(self):
        """Make this solver's partition the one the C library's collectives use (several solvers,
        e.g. one per load-case mesh, may share the device context)."""
        from ._lib import lib, check
        if self.world == 1:
            return
        p = self.plan
        keep = [np.ascontiguousarray(a, dtype=np.int64) for a in (p.offsets, p.need_lo, p.need_hi, p.give_lo, p.give_hi)]
        check(self.ctx.h, lib.myc_dist_set_plan(self.ctx.h, *[a.ctypes.data_as(C.c_void_p) for a in keep]))
        self.ctx.node_offsets = p.offsets

    def _ensure_peer(self, n_cols):
        """Set up (or grow) the NVLink peer-memory buffers of the fused multi-GPU PCG.  Collective.
        If any rank cannot open a peer handle every rank falls back to the NCCL loop."""
        import os
        import torch
        import torch.distributed as dist
        from ._lib import lib, check
        if self.world == 1 or os.environ.get("MYC_NO_PEER") == "1":
            return
        if getattr(self.ctx, "peer_cap", 0) >= n_cols:
            return
        cap = int(n_cols * 1.05) + 1024
        handle = np.zeros(64, dtype=np.uint8)
        if getattr(self.ctx, "peer_cap", 0) > 0:      # growing: importers unmap before the exporters free (CUDA IPC rule)
            torch.cuda.synchronize()
            dist.barrier()
            lib.myc_dist_release_peers(self.ctx.h)
            dist.barrier()
        check(self.ctx.h, lib.myc_dist_peer_alloc(self.ctx.h, cap, handle.ctypes.data_as(C.c_void_p)))
        t = torch.from_numpy(handle).to(self.ctx.device)
        gathered = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(gathered, t)
        allh = np.ascontiguousarray(np.concatenate([g.cpu().numpy() for g in gathered]))
        rc = lib.myc_dist_peer_open(self.ctx.h, allh.ctypes.data_as(C.c_void_p))
        ok = torch.tensor([1 if rc == 0 else 0], device=self.ctx.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            lib.myc_dist_peer_disable(self.ctx.h)
            self.ctx.peer_cap = 1 << 62          # do not retry
            self.ctx.peer_enabled = False
        else:
            self.ctx.peer_cap = cap
            self.ctx.peer_enabled = True

    def assemble(self, E, A, I):
        from . import device as dv
        self._install_plan()
        return dv.assemble(self.ctx, self.mesh, E, A, I, node_range=(self.plan.node_begin, self.plan.node_end))

    def true_residual(self, K, system, x):
        """||b - A x|| / ||b|| recomputed from scratch with THIS solver's halo plan (collective)."""
        from . import device as dv
        self._install_plan()
        return dv.true_residual(self.ctx, K, system, x)

    def load_case(self, K, known_dofs, known_vals, react_dofs=None, rtol=1e-10, precond="amg",
                  maxit=500_000, reg=1e-12, gather_U=True, system=None, x0=None):
        """Returns dict(U (global, on every rank if gather_U), iterations, relres, total_force).
        ``precond``: "amg" (aggregation multigrid, built collectively; prepared as "block6" where the hierarchy is
        not applicable), "block6", "block3" or "jacobi".  ``system``: reuse the Dirichlet system (and its
        preconditioner) of an earlier call for the same K and known DOFs -- only the prescribed values are new.
        ``x0``: starting guess on this rank's rows (overwritten with the solution)."""
        import torch
        from . import device as dv
        from ._lib import lib, check
        ctx = self.ctx
        dev = ctx.device
        self._install_plan()
        td = lambda a, dt: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        kd, kv = td(known_dofs, np.int64), td(known_vals, np.float64)
        if precond == "block12" or (precond == "block6" and not (self.block6 and K.row_offset % 6 == 0)):
            # the 12-row groups would need cuts on multiples of 4 nodes: a row-partitioned solve uses the
            # 3x3 node blocks instead
            precond = "block3"
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        sysd = dv.apply_dirichlet(ctx, K, kd, kv, reg, precond=precond, reuse=system)
        precond = sysd.precond
        if precond != "amg":
            self._ensure_peer(K.n_cols)
        ev[1].record()
        x, iters, relres = dv.pcg(ctx, K, sysd, x0=x0, precond=precond, rtol=rtol, maxit=maxit)
        U = dv.merge_solution(ctx, K, sysd, x)         # own rows of a zeroed global vector
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        total_force = None
        if self.world > 1:
            if gather_U:
                check(ctx.h, lib.myc_allgather_owned(ctx.h, C.c_void_p(U.data_ptr()), stream))
            elif react_dofs is not None:               # K @ U on the own rows only needs the neighbours' boundary values
                check(ctx.h, lib.myc_halo_exchange(ctx.h, C.c_void_p(U.data_ptr()), stream))
        if react_dofs is not None:
            rd = np.asarray(react_dofs, dtype=np.int64)
            lo, hi = K.row_offset, K.row_offset + K.n_rows
            mine = rd[(rd >= lo) & (rd < hi)] - lo
            F = dv.spmv(ctx, K, U)                      # local rows of K @ U   (fea_solver.py:257)
            part = dv.gather_sum(ctx, F, td(mine, np.int64)) if len(mine) else 0.0
            buf = (C.c_double * 1)(part)
            check(ctx.h, lib.myc_allreduce_sum(ctx.h, buf, 1, stream))
            total_force = float(buf[0])
        ev[2].record()
        ev[2].synchronize()
        return {"U": U, "x": x, "system": sysd, "iterations": iters, "relres": relres, "total_force": total_force,
                "precond": precond, "ms_setup": ev[0].elapsed_time(ev[1]), "ms_solve": ev[1].elapsed_time(ev[2])}
