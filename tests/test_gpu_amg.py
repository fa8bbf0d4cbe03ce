"""Aggregation-multigrid PCG (MYC_PC_AMG) against the numpy restatement (oracle/amg_oracle.py) and the
reference's direct solve (oracle.fea_oracle.solve_system = src/fea_solver.py:112-135)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from mycelium_fea_project_b200 import device as dv, fea_solver as fs  # noqa: E402
from mycelium_fea_project_b200.synth import synth_network  # noqa: E402
from oracle import fea_oracle as fo, amg_oracle as ao  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return dv.Context.get()


def _problem(N, case="Y", tol=0.5, seed=0):
    coords, n1, n2 = synth_network(N, seed=seed)
    axis, comp = fs.LOAD_CASES[case]
    hi, lo = fs.grip_nodes(coords, tol, axis)
    kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, comp)
    return coords, n1, n2, kd, kv, hi, comp


def _prepare(ctx, coords, n1, n2, kd, kv, active=None):
    mesh = dv.DeviceMesh.from_host(coords, n1, n2, active)
    K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
    sysd = dv.apply_dirichlet(ctx, K, torch.from_numpy(kd).cuda(), torch.from_numpy(kv).cuda(), precond="amg")
    return K, sysd


@pytest.mark.parametrize("N,case", [(64, "Y"), (96, "X"), (128, "Y")])
def test_hierarchy_matches_numpy_restatement(ctx, N, case):
    """Aggregates of every level and the level sizes equal the numpy algorithm run on the SAME K
    (the GPU's CSR is downloaded, so strengths are bitwise the same numbers)."""
    coords, n1, n2, kd, kv, _, _ = _problem(N, case)
    K, sysd = _prepare(ctx, coords, n1, n2, kd, kv)
    assert sysd.precond == "amg" and sysd.amg_levels >= 2
    levels, setup_ms = dv.amg_levels(ctx)
    free = np.ones(K.n_rows, bool)
    free[kd] = False
    L0 = ao.level_from_csr(K.to_scipy(), free)
    ref = ao.build_hierarchy(L0)
    assert [l[0] for l in levels] == [l.n for l in ref]
    assert [l[1] for l in levels[1:]] == [len(l.bnode) for l in ref[1:]]      # level 0 also stores blocks of known nodes
    for l in range(len(ref) - 1):
        agg = dv.amg_aggregates(ctx, l).cpu().numpy()
        assert np.array_equal(agg, ref[l].agg), f"aggregates differ on level {l}"


@pytest.mark.parametrize("N,case", [(64, "Y"), (128, "X"), (256, "Y")])
def test_amg_pcg_matches_direct_solve_and_numpy_iterations(ctx, N, case):
    coords, n1, n2, kd, kv, hi, comp = _problem(N, case, tol=1.5 if N >= 128 else 0.5)
    K, sysd = _prepare(ctx, coords, n1, n2, kd, kv)
    x, iters, relres = dv.pcg(ctx, K, sysd, precond="amg", rtol=1e-12)
    U = dv.merge_solution(ctx, K, sysd, x).cpu().numpy()
    Ko = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
    Uo = fo.solve_system(Ko, kd, kv)
    err = np.linalg.norm(U - Uo) / np.linalg.norm(Uo)
    assert err <= 1e-8, err
    assert np.array_equal(U[kd], kv)
    assert dv.true_residual(ctx, K, sysd, x) <= 1e-11
    # iteration count: numpy PCG with the same hierarchy algorithm (standard recurrence; the GPU runs the
    # single-reduction form, so allow a couple of iterations of slack)
    free = np.ones(K.n_rows, bool)
    free[kd] = False
    _, it_ref, _ = ao.amg_pcg(K.to_scipy(), free, sysd.rhs.cpu().numpy(), rtol=1e-12)
    assert abs(iters - it_ref) <= max(2, it_ref // 20), (iters, it_ref)
    # and far fewer iterations than block-Jacobi needs
    s6 = dv.apply_dirichlet(ctx, K, torch.from_numpy(kd).cuda(), torch.from_numpy(kv).cuda(), precond="block6")
    _, it6, _ = dv.pcg(ctx, K, s6, precond="block6", rtol=1e-12)
    assert iters * 4 < it6, (iters, it6)


def test_amg_with_failed_elements_and_second_rhs(ctx):
    """Deactivated elements (floating pieces, isolated nodes) and a second right-hand side on the same hierarchy."""
    coords, n1, n2, kd, kv, hi, comp = _problem(96, "Y")
    act = np.random.default_rng(5).random(len(n1)) > 0.25
    K, sysd = _prepare(ctx, coords, n1, n2, kd, kv, act)
    Ko = fo.assemble_global_stiffness(coords, n1, n2, act)
    for scale in (1.0, -2.5):
        sysd.rhs.mul_(scale) if scale != 1.0 else None
        x, iters, _ = dv.pcg(ctx, K, sysd, precond="amg", rtol=1e-12)
        U = dv.merge_solution(ctx, K, sysd, x).cpu().numpy()
        if scale == 1.0:
            Uo = fo.solve_system(Ko, kd, kv)
            assert np.linalg.norm(U - Uo) / np.linalg.norm(Uo) <= 1e-8
            x1 = x.clone()
        else:
            assert float(torch.linalg.norm(x - scale * x1) / torch.linalg.norm(x1)) <= 1e-9


def test_amg_falls_back_for_partial_node_constraints(ctx):
    """A Dirichlet set that prescribes only one DOF of a node: the system is prepared for block-Jacobi instead."""
    coords, n1, n2, kd, kv, _, _ = _problem(64)
    kd2, kv2 = kd[kd % 3 != 2], kv[kd % 3 != 2]          # z left free on the grips
    mesh = dv.DeviceMesh.from_host(coords, n1, n2)
    K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
    sysd = dv.apply_dirichlet(ctx, K, torch.from_numpy(kd2).cuda(), torch.from_numpy(kv2).cuda(), precond="amg")
    assert sysd.precond == "block6" and sysd.amg_levels == 0
    x, iters, _ = dv.pcg(ctx, K, sysd, precond="amg", rtol=1e-12)
    U = dv.merge_solution(ctx, K, sysd, x).cpu().numpy()
    Uo = fo.solve_system(fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool)), kd2, kv2)
    assert np.linalg.norm(U - Uo) / np.linalg.norm(Uo) <= 1e-8
