"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): the row-partitioned NCCL solve
must reproduce the single-GPU solve and the oracle's direct solve, rank-local CSR blocks must be
verbatim slices of the global CSR."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs >= 2 CUDA devices", allow_module_level=True)

import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, N, precond, ret, no_peer=False, shape=None, case="Y", env=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    if no_peer:
        os.environ["MYC_NO_PEER"] = "1"
    os.environ.update(env or {})
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mycelium_fea_project_b200 import device as dv, dist as md, fea_solver as fs
        from mycelium_fea_project_b200.synth import synth_network
        from oracle import fea_oracle as fo
        coords, n1, n2 = synth_network(N) if shape is None else synth_network(*shape)
        axis, comp = fs.LOAD_CASES[case]
        solver = md.DistributedSolver((coords, n1, n2), device=torch.device("cuda", rank))
        K = solver.assemble(fs.E_mod, fs.A, fs.I)
        Ko = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
        lo, hi = K.row_offset, K.row_offset + K.n_rows
        Ks = K.to_scipy()
        ref = Ko[lo:hi]
        assert np.array_equal(Ks.indptr, ref.indptr) and np.array_equal(Ks.indices, ref.indices)
        assert np.abs(Ks.data - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
        hi_n, lo_n = fs.grip_nodes(coords, 0.5, axis)
        kd, kv = fs.build_bc(hi_n, lo_n, 0.02, -0.02, comp)
        out = solver.load_case(K, kd, kv, react_dofs=3 * hi_n + comp, rtol=1e-12, precond=precond)
        U = out["U"].cpu().numpy()
        Uo = fo.solve_system(Ko, kd, kv)
        err = np.linalg.norm(U - Uo) / np.linalg.norm(Uo)
        tf = (Ko @ Uo)[3 * hi_n + comp].sum()
        assert err <= 1e-8, err
        assert abs(out["total_force"] - tf) <= 1e-7 * abs(tf)
        tr = dv.true_residual(solver.ctx, K, out["system"], out["x"])
        assert tr <= 1e-11
        # a second load case on the same solver (epochs of the peer flags continue across solves)
        kd2, kv2 = fs.build_bc(hi_n, lo_n, 0.01, -0.03, comp)
        out2 = solver.load_case(K, kd2, kv2, rtol=1e-12, precond=precond)
        U2o = fo.solve_system(Ko, kd2, kv2)
        err2 = np.linalg.norm(out2["U"].cpu().numpy() - U2o) / np.linalg.norm(U2o)
        assert err2 <= 1e-8, err2
        extra = None
        if precond == "amg":
            assert out["system"].precond == "amg", "the multigrid hierarchy was not built"
            levels, _ = dv.amg_levels(solver.ctx, detail=True)
            extra = [(l["n_global"], l["replicated"]) for l in levels]
            if rank == 0:
                # the numpy restatement with the same row partition and replication threshold: same level sizes,
                # same iteration count (it runs the standard recurrence, the GPU the single-reduction form)
                from oracle import amg_oracle as ao
                ao.REPLICATE_NODES = int(os.environ.get("MYC_AMG_REPLICATE_NODES", ao.REPLICATE_NODES))
                free = np.ones(Ko.shape[0], bool)
                free[kd] = False
                ubc = np.zeros(Ko.shape[0])
                ubc[kd] = kv
                b = -(Ko @ ubc)
                b[kd] = 0.0
                _, it_ref, lv_ref = ao.amg_pcg(Ko, free, b, rtol=1e-12, node_offsets=solver.plan.offsets)
                assert [l[0] for l in extra] == [l.n for l in lv_ref], (extra, [l.n for l in lv_ref])
                assert abs(out["iterations"] - it_ref) <= max(2, it_ref // 20), (out["iterations"], it_ref)
        ret[rank] = (out["iterations"], err, out["total_force"], bool(getattr(solver.ctx, "peer_enabled", False)), extra)
        md.shutdown(solver.ctx)          # orderly: unmap peers' buffers, barrier, free
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precond,no_peer", [("jacobi", False), ("jacobi", True), ("block3", False)])
def test_two_gpu_solve_matches_oracle(precond, no_peer):
    """no_peer False: persistent solver kernel over NVLink peer memory (jacobi and block3); True: NCCL loop."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 96, precond, ret, no_peer), nprocs=world, join=True)
    assert len(ret) == world
    assert ret[0][0] == ret[1][0]            # same iteration count on both ranks
    assert ret[0][2] == ret[1][2]            # identical all-reduced reaction
    if precond == "jacobi" and not no_peer:
        assert ret[0][3] and ret[1][3], "peer-memory path was not enabled on this box"


def test_two_gpu_block6():
    """Cuts on even nodes, block6 inside the peer-memory solver kernel: same parity bars, and fewer
    iterations than the 3x3 node blocks need on the same two-rank problem."""
    world = 2
    mgr = mp.Manager()
    ret6, ret3 = mgr.dict(), mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 96, "block6", ret6, False), nprocs=world, join=True)
    mp.spawn(_worker, args=(world, _free_port(), 96, "block3", ret3, False), nprocs=world, join=True)
    assert len(ret6) == world and ret6[0][0] == ret6[1][0] and ret6[0][2] == ret6[1][2]
    assert ret6[0][0] < ret3[0][0], (ret6[0][0], ret3[0][0])


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs >= 4 CUDA devices")
@pytest.mark.parametrize("case,no_peer", [("X", False), ("Y", False), ("X", True)])
def test_four_gpu_strips(case, no_peer):
    """Middle ranks have two neighbours; X: tall specimen with grips (known DOFs) on every rank,
    Y: wide specimen with grips on the first and last rank only (the bench's weak-scaling shapes)."""
    world = 4
    shape = (4 * 40, 48) if case == "X" else (48, 4 * 40)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 0, "jacobi", ret, no_peer, shape, case), nprocs=world, join=True)
    assert len(ret) == world
    assert len({v[0] for v in ret.values()}) == 1 and len({v[2] for v in ret.values()}) == 1


@pytest.mark.parametrize("replicate_nodes", [0, 300, 65536])
def test_two_gpu_amg(replicate_nodes):
    """Aggregation-multigrid PCG on a row-partitioned mesh: partitioned levels exchange their correction vectors
    through NVLink peer memory, levels at or below the replication threshold are processed by every rank in
    full.  0: every level partitioned; 300: two partitioned coarse levels, then the seam; 65536 (default): the
    first coarse level is already replicated.  Same parity bars as every solver (U vs the direct solve <= 1e-8),
    level sizes and iteration count equal to the numpy restatement with the same partition."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 128, "amg", ret, False, None, "Y",
                            {"MYC_AMG_REPLICATE_NODES": str(replicate_nodes)}), nprocs=world, join=True)
    assert len(ret) == world
    assert ret[0][0] == ret[1][0] and ret[0][2] == ret[1][2]
    assert ret[0][4] == ret[1][4]                       # same hierarchy shape on both ranks
    kinds = [k for _, k in ret[0][4]]
    if replicate_nodes == 0:
        assert set(kinds) == {0}
    else:
        assert kinds[0] == 0 and 1 in kinds and kinds[-1] in (1, 2)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs >= 4 CUDA devices")
@pytest.mark.parametrize("case,replicate_nodes", [("X", 300), ("Y", 300), ("Y", 65536)])
def test_four_gpu_amg(case, replicate_nodes):
    world = 4
    shape = (4 * 48, 64) if case == "X" else (64, 4 * 48)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 0, "amg", ret, False, shape, case,
                            {"MYC_AMG_REPLICATE_NODES": str(replicate_nodes)}), nprocs=world, join=True)
    assert len(ret) == world
    assert len({v[0] for v in ret.values()}) == 1 and len({v[2] for v in ret.values()}) == 1
    assert len({tuple(v[4]) for v in ret.values()}) == 1


def _ramp_worker(rank, world, port, golden_dir, ret):
    import gzip
    import io
    import pandas as pd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mycelium_fea_project_b200 import device as dv, dist as md, fea_solver as fs
        d = os.path.join(golden_dir, "ref_results", "sim_20251117_181147")
        nodes = pd.read_csv(io.BytesIO(gzip.open(os.path.join(d, "nodes.csv.gz")).read()))
        elems = pd.read_csv(io.BytesIO(gzip.open(os.path.join(d, "elements.csv.gz")).read()))
        g = np.load(os.path.join(d, "fea_results", "active_elements.npz"))
        gold = np.unpackbits(g["packed"], axis=1)[:, :int(g["n_elems"])].astype(bool)
        fd = pd.read_csv(os.path.join(d, "fea_results", "force_displacement.csv"), float_precision="round_trip").values
        rec = fs.fea_ramp_distributed(nodes[["x", "y", "z"]].values, elems["n1"].values, elems["n2"].values)
        act = np.array(rec["active"])
        assert act.shape == gold.shape and np.array_equal(act, gold), "failure cascade differs from the reference's"
        got = np.array(rec["force_disp"])
        assert np.array_equal(got[:, 0], fd[:, 0])
        assert np.abs(got[:, 1] - fd[:, 1]).max() <= 1e-7 * np.abs(fd[:, 1]).max()
        assert not all(rec["reassembled"]) and rec["reassembled"][0]
        ret[rank] = (rec["iterations"], float(np.abs(np.array(rec["stress"])).sum()), float(np.abs(np.array(rec["disp"])).sum()))
        md.shutdown(dv.Context.get())
    finally:
        dist.destroy_process_group()


def test_two_gpu_ramp_matches_reference_cascade(golden_dir):
    """The 40-step displacement ramp of results/sim_20251117_181147 row-partitioned over two GPUs (fea_ramp_distributed:
    incremental re-assembly, warm start, multigrid PCG, strain / failure on each rank's elements): the failure
    cascade equals the reference's committed active_elements.csv, the force curve agrees to 1e-7, and both ranks hold
    identical records."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ramp_worker, args=(world, _free_port(), golden_dir, ret), nprocs=world, join=True)
    assert len(ret) == world and ret[0] == ret[1]
