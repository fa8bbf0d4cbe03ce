"""world_size-2/3 CPU tests (gloo) of the multi-GPU host logic: partition, halo-range plan, the
halo exchange protocol and the row-partitioned PCG recurrences -- the same protocol
csrc/dist.cu + csrc/pcg.cu run on NCCL.  The numerics here are numpy/scipy (the oracle's K);
no CUDA is touched."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mycelium_fea_project_b200 import dist as md
from mycelium_fea_project_b200.synth import synth_network
from oracle import fea_oracle as fo


def test_partition_and_plan_single_process():
    coords, n1, n2 = synth_network(64)
    n = len(coords)
    for world in (1, 2, 3, 8):
        off = md.partition_nodes(n, world)
        assert off[0] == 0 and off[-1] == n and len(off) == world + 1
        assert np.all(np.diff(off) >= n // world) and np.all(np.diff(off) <= n // world + 1)
        K = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
        for r in range(world):
            p = md.make_plan(n1, n2, None, n, r, world)
            rows = K[3 * off[r]:3 * off[r + 1]]
            cols = np.unique(rows.indices) // 3
            ext = cols[(cols < off[r]) | (cols >= off[r + 1])]
            # every external column node lies inside the range requested from its owner
            own = md.owner_of(ext, off)
            for q in np.unique(own):
                e = ext[own == q]
                assert p.need_lo[q] <= e.min() and e.max() < p.need_hi[q]
                assert off[q] <= p.need_lo[q] and p.need_hi[q] <= off[q + 1]
            # strip partition of a row-major grid: only the two neighbours are needed
            assert set(np.nonzero(p.need_hi > p.need_lo)[0]) <= {r - 1, r + 1}
            # give is the transpose of need
            for q in range(world):
                pq = md.make_plan(n1, n2, None, n, q, world)
                assert p.give_lo[q] == pq.need_lo[r] and p.give_hi[q] == pq.need_hi[r]


def test_aligned_partition_keeps_node_pairs_on_one_rank():
    """The multi-GPU partition (cuts aligned to the 6x6 Jacobi blocks): every interior cut on an even node (or at the end), still balanced."""
    for n, world in [(10, 3), (11, 4), (174738, 8), (7, 8), (2, 4), (0, 2), (1067, 2), (1, 1)]:
        off = md.partition_nodes(n, world, align=2)
        assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) >= 0)
        assert np.all((off[1:-1] % 2 == 0) | (off[1:-1] == n))
        if n >= 4 * world:
            assert np.diff(off).max() - np.diff(off).min() <= 2
        assert np.array_equal(md.partition_nodes(n, world, align=1), md.partition_nodes(n, world))
    coords, n1, n2 = synth_network(24)
    n = len(coords)
    for r in range(3):
        p = md.make_plan(n1, n2, None, n, r, 3, align=2)
        assert p.node_begin % 2 == 0
        # halo ranges still cover every remote neighbour of the rank's own nodes
        b, e = p.node_begin, p.node_end
        far = np.concatenate([n2[(n1 >= b) & (n1 < e)], n1[(n2 >= b) & (n2 < e)]])
        far = far[(far < b) | (far >= e)]
        own = md.owner_of(far, p.offsets)
        for q in np.unique(own):
            f = far[own == q]
            assert p.need_lo[q] <= f.min() and f.max() < p.need_hi[q]


def test_locality_order_makes_ranges_small():
    coords, n1, n2 = synth_network(32)
    rng = np.random.default_rng(0)
    shuffle = rng.permutation(len(coords))              # destroy the numbering
    inv = np.empty_like(shuffle); inv[shuffle] = np.arange(len(shuffle))
    c2, a2, b2 = coords[shuffle], inv[n1], inv[n2]
    off = md.partition_nodes(len(c2), 4)
    lo, hi = md.halo_ranges(a2, b2, None, off, 1)
    assert (hi - lo).sum() > 0.5 * len(c2)              # all-to-all without reordering
    perm = md.locality_order(c2, axis=1)
    inv2 = np.empty_like(perm); inv2[perm] = np.arange(len(perm))
    lo, hi = md.halo_ranges(inv2[a2], inv2[b2], None, off, 1)
    assert (hi - lo).sum() < 0.15 * len(c2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, N, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coords, n1, n2 = synth_network(N)
        n = len(coords)

        def all_gather(arr):
            out = [None] * world
            dist.all_gather_object(out, arr)
            return out

        plan = md.make_plan(n1, n2, None, n, rank, world, all_gather)
        plan_local = md.make_plan(n1, n2, None, n, rank, world)
        assert np.array_equal(plan.give_lo, plan_local.give_lo) and np.array_equal(plan.give_hi, plan_local.give_hi)
        K = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
        lo, hi = 3 * plan.node_begin, 3 * plan.node_end
        Kloc = K[lo:hi]
        # --- halo exchange: own slice valid, rest poisoned, then refreshed from the owners
        xfull = np.random.default_rng(1).standard_normal(3 * n)
        xg = torch.full((3 * n,), float("nan"), dtype=torch.float64)
        xg[lo:hi] = torch.from_numpy(xfull[lo:hi])
        md.exchange_halo_torch(xg, plan)
        y = Kloc @ np.nan_to_num(xg.numpy(), nan=1e300)
        assert np.array_equal(y, (K @ xfull)[lo:hi]), "halo did not cover every external column"
        # --- row-partitioned Jacobi-PCG with the recurrences of csrc/pcg.cu
        top, bot = fo.grip_nodes(coords, 0.5)
        kd, kv = fo.build_bc(top, bot, 0.02, -0.02)
        ubc = np.zeros(3 * n); ubc[kd] = kv
        free = np.ones(3 * n, bool); free[kd] = False
        fl = free[lo:hi]
        b = np.where(fl, -(Kloc @ ubc), 0.0)
        dinv = np.where(fl, 1.0 / (Kloc.diagonal(k=lo) + 1e-12), 0.0)

        def allsum(*v):
            t = torch.tensor(v, dtype=torch.float64)
            dist.all_reduce(t)
            return t.tolist()

        x = np.zeros(hi - lo); r = b.copy()
        pg = torch.zeros(3 * n, dtype=torch.float64)
        pg[lo:hi] = torch.from_numpy(dinv * r)
        (rz, bb) = allsum(float(r @ (dinv * r)), float(b @ b))
        it = 0
        while True:
            md.exchange_halo_torch(pg, plan)
            p = pg.numpy()
            Ap = Kloc @ p + 1e-12 * p[lo:hi]
            (pAp,) = allsum(float(p[lo:hi] @ Ap))
            alpha = rz / pAp
            x += alpha * p[lo:hi]
            r = np.where(fl, r - alpha * Ap, 0.0)
            it += 1
            rz_new, rr = allsum(float(r @ (dinv * r)), float(r @ r))
            if rr <= (1e-12 ** 2) * bb or it > 5000:
                break
            pg[lo:hi] = torch.from_numpy(dinv * r + (rz_new / rz) * p[lo:hi])
            rz = rz_new
        U = ubc.copy()
        Uloc = np.where(fl, x, ubc[lo:hi])
        gathered = all_gather(Uloc)
        U = np.concatenate(gathered)
        Uo = fo.solve_system(K, kd, kv)
        err = np.linalg.norm(U - Uo) / np.linalg.norm(Uo)
        assert err < 1e-8, err
        ret[rank] = (it, err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_partitioned_pcg_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 48, ret), nprocs=world, join=True)
    assert len(ret) == world
    its = {v[0] for v in ret.values()}
    assert len(its) == 1          # every rank took the same number of iterations


@pytest.mark.parametrize("world", [2, 3])
def test_local_mesh_slice_is_what_a_rank_assembles_from(world):
    """The per-rank share of the mesh (DistributedSolver uploads only this): assembling a rank's rows from its
    incident elements alone gives the verbatim row slice of the global K, and the coordinate range covers every
    end node of those elements."""
    coords, n1, n2 = synth_network(24, 40, seed=5)
    active = np.ones(len(n1), bool)
    K = fo.assemble_global_stiffness(coords, n1, n2, active)
    off = md.partition_nodes(len(coords), world, align=2)
    seen = np.zeros(len(n1), int)
    for r in range(world):
        idx, lo, hi = md.local_mesh_slice(n1, n2, off[r], off[r + 1])
        seen[idx] += 1
        assert np.all(np.diff(idx) > 0)                                  # element order kept
        assert lo <= off[r] and hi >= off[r + 1]
        assert min(n1[idx].min(), n2[idx].min()) >= lo and max(n1[idx].max(), n2[idx].max()) < hi
        c = np.full_like(coords, np.nan)
        c[lo:hi] = coords[lo:hi]                                         # nothing outside the range may be read
        Kr = fo.assemble_global_stiffness(c, n1[idx], n2[idx], np.ones(len(idx), bool))[3 * off[r]:3 * off[r + 1]]
        ref = K[3 * off[r]:3 * off[r + 1]]
        assert np.array_equal(Kr.indptr, ref.indptr) and np.array_equal(Kr.indices, ref.indices)
        assert np.array_equal(Kr.data, ref.data)
    assert seen.min() >= 1 and seen.max() <= 2                           # cut elements live on both sides


@pytest.mark.parametrize("world,replicate", [(1, 0), (2, 0), (4, 0), (4, 500), (3, 10 ** 9)])
def test_partitioned_multigrid_restatement_solves_the_system(world, replicate, monkeypatch):
    """oracle/amg_oracle.py with a row partition (aggregates confined to a rank above the replication threshold,
    unrestricted below it) is still a symmetric positive preconditioner: PCG reaches the direct solve, in a number
    of iterations close to the unpartitioned hierarchy's.  (The GPU tests compare the CUDA path with this.)"""
    from oracle import amg_oracle as ao
    monkeypatch.setattr(ao, "REPLICATE_NODES", replicate)
    coords, n1, n2 = synth_network(64, seed=0)
    K = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
    hi, lo = fo.grip_nodes(coords, 0.5)
    kd, kv = fo.build_bc(hi, lo, 0.02, -0.02)
    free = np.ones(K.shape[0], bool)
    free[kd] = False
    ubc = np.zeros(K.shape[0])
    ubc[kd] = kv
    b = -(K @ ubc)
    b[kd] = 0.0
    off = md.partition_nodes(len(coords), world, 2) if world > 1 else None
    x, it, levels = ao.amg_pcg(K, free, b, rtol=1e-12, node_offsets=off)
    x0, it0, levels0 = ao.amg_pcg(K, free, b, rtol=1e-12)
    U = x + ubc
    Uo = fo.solve_system(K, kd, kv)
    assert np.linalg.norm(U - Uo) <= 1e-8 * np.linalg.norm(Uo)
    assert it <= 2 * it0 + 5
    for l in levels[:-1]:                                   # aggregates never span ranks while the level is partitioned
        own = getattr(l, "owner", None)
        if own is not None:
            m = l.agg >= 0
            first = np.full(l.n_coarse, -1)
            first[l.agg[m][::-1]] = own[m][::-1]
            assert np.array_equal(first[l.agg[m]], own[m])
    if world > 1 and replicate >= 10 ** 9:
        assert not hasattr(levels[1], "owner")              # replicated from the first coarse level on
