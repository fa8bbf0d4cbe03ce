"""Size-independent properties of the CUDA path at a BASELINE.json full size (configs[2]: the
synthetic 2048x2048 network, 8.39 M DOF), where the oracle's direct solve no longer finishes in
seconds.  Everything here follows from the reference's formulas alone:

* K_e = [[S,-S],[-S,S]] (src/fea_solver.py:60-66)  =>  every rigid translation is in the null
  space of the assembled K (K t = 0) and K is symmetric (x.Ky = y.Kx);
* scipy's CSR (src/fea_solver.py:105): row_ptr monotone from 0 to nnz, columns strictly ascending
  within a row (duplicates merged), three equal-length rows per node;
* solve_system (src/fea_solver.py:131-133): U[known] == known_vals exactly, and with K t_y = 0 the
  y-reactions of the two grips cancel up to the solver's residual and the 1e-12 I shift
  (src/fea_solver.py:125):  R_top + R_bot - 1e-12 sum_free U_y = sum_free (b - A x)_y, which is
  bounded by sqrt(n) ||b - A x||  (identity checked on the CPU oracle at 128^2).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from mycelium_fea_project_b200 import device as dv  # noqa: E402
from mycelium_fea_project_b200 import fea_solver as fs  # noqa: E402
from mycelium_fea_project_b200.synth import synth_network  # noqa: E402

N = 2048


@pytest.fixture(scope="module")
def big():
    ctx = dv.Context.get()
    coords, n1, n2 = synth_network(N)
    mesh = dv.DeviceMesh.from_host(coords, n1, n2)
    K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
    return ctx, coords, n1, n2, mesh, K


def test_counts_match_the_generator_contract(big):
    """0.667 N^2 nodes, 0.887 N^2 elements, ~10.95 nnz/row; the seeded
    generator makes the counts exact and the assembly must reproduce them on every box."""
    ctx, coords, n1, n2, mesh, K = big
    assert K.n_rows == 3 * len(coords) == 8_388_996
    assert K.nnz == 91_943_955
    assert abs(len(n1) / N ** 2 - 0.887) < 5e-3 and abs(K.nnz / K.n_rows - 10.95) < 0.02


def test_csr_structure_invariants(big):
    ctx, coords, n1, n2, mesh, K = big
    rp = K.row_ptr.long()
    assert int(rp[0]) == 0 and int(rp[-1]) == K.nnz
    lens = rp[1:] - rp[:-1]
    assert int(lens.min()) >= 0
    # three equal-length rows per node, each a whole number of 3-wide node blocks
    l3 = lens.view(-1, 3)
    assert bool((l3[:, 0] == l3[:, 1]).all()) and bool((l3[:, 1] == l3[:, 2]).all())
    assert bool((lens % 3 == 0).all())
    # columns strictly ascending inside every row: a decrease is only allowed at a row boundary
    ci = K.col_idx
    not_ascending = (ci[1:] <= ci[:-1]).nonzero().flatten() + 1      # positions that start a new run
    is_row_start = torch.zeros(K.nnz + 1, dtype=torch.bool, device=ci.device)
    is_row_start[rp] = True
    assert bool(is_row_start[not_ascending].all()), "columns not strictly ascending inside a row"
    assert int(ci.min()) >= 0 and int(ci.max()) < K.n_cols
    # rows of isolated sites are empty (about 1.2 % of the nodes), nothing else is
    deg = np.bincount(np.concatenate([n1, n2]), minlength=len(coords))
    empty_nodes = torch.from_numpy(deg == 0).to(ci.device)
    assert bool(((l3[:, 0] == 0) == empty_nodes).all())
    assert 0.005 < float(empty_nodes.double().mean()) < 0.02
    # nnz follows from the mesh alone: 9 per node with an element + 9 per distinct directed pair
    a, b = np.minimum(n1, n2).astype(np.int64), np.maximum(n1, n2).astype(np.int64)
    pairs = np.unique(a * len(coords) + b).size
    assert K.nnz == 9 * int((deg > 0).sum()) + 18 * pairs


def test_assembly_is_bit_reproducible(big):
    ctx, coords, n1, n2, mesh, K = big
    K2 = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
    assert torch.equal(K.row_ptr, K2.row_ptr) and torch.equal(K.col_idx, K2.col_idx) and torch.equal(K.val, K2.val)


def test_rigid_translations_are_in_the_null_space(big):
    ctx, coords, n1, n2, mesh, K = big
    rp = K.row_ptr.long()
    rows = torch.repeat_interleave(torch.arange(K.n_rows, device=K.val.device), rp[1:] - rp[:-1])
    rowmax = torch.zeros(K.n_rows, dtype=torch.float64, device=K.val.device)
    rowmax.scatter_reduce_(0, rows, K.val.abs(), reduce="amax", include_self=True)
    for comp in range(3):
        t = torch.zeros(K.n_rows, dtype=torch.float64, device=K.val.device)
        t[comp::3] = 1.0
        y = dv.spmv(ctx, K, t)
        # a row sums <= ~15 entries of alternating sign: a few ulp of the largest one
        assert bool((y.abs() <= 64 * np.finfo(float).eps * rowmax).all()), f"K t != 0 for translation {comp}"


def test_operator_is_symmetric_and_linear(big):
    ctx, coords, n1, n2, mesh, K = big
    g = torch.Generator(device=K.val.device).manual_seed(1)
    x = torch.randn(K.n_rows, dtype=torch.float64, device=K.val.device, generator=g)
    y = torch.randn(K.n_rows, dtype=torch.float64, device=K.val.device, generator=g)
    Kx, Ky = dv.spmv(ctx, K, x), dv.spmv(ctx, K, y)
    scale = float(Kx.norm() * y.norm())
    assert abs(float(torch.dot(y, Kx) - torch.dot(x, Ky))) <= 1e-12 * scale
    z = dv.spmv(ctx, K, 2.0 * x - 0.5 * y)
    assert float((z - (2.0 * Kx - 0.5 * Ky)).norm()) <= 1e-13 * float(z.norm())
    # positive semi-definite: x.Kx = sum_e (dx_e . S_e dx_e) >= 0
    assert float(torch.dot(x, Kx)) > 0.0


def test_full_solve_residual_and_grip_equilibrium(big):
    """One complete Y load case at 8.39 M DOF: true residual at the requested tolerance, prescribed
    DOFs exact, grip reactions in equilibrium."""
    ctx, coords, n1, n2, mesh, K = big
    hi, lo = fs.grip_nodes(coords, 1.5, 1)
    kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, 1)
    res = fs.analyze_load_case(mesh, kd, kv, react_dofs=3 * hi + 1, rtol=1e-10, K=K)
    tr = dv.true_residual(ctx, K, res.system, res.x)
    assert res.relres <= 1e-10 and tr <= 1.2e-10
    dev = res.U.device
    kd_d = torch.from_numpy(kd).to(dev)
    assert torch.equal(res.U[kd_d], torch.from_numpy(kv).to(dev)), "U[known] must equal known_vals exactly"
    F = dv.spmv(ctx, K, res.U)
    r_top = dv.gather_sum(ctx, F, torch.from_numpy((3 * hi + 1).astype(np.int64)).to(dev))
    r_bot = dv.gather_sum(ctx, F, torch.from_numpy((3 * lo + 1).astype(np.int64)).to(dev))
    assert abs(r_top - res.total_force) <= 1e-12 * abs(r_top)
    assert r_top > 0.0 > r_bot                               # the stretched specimen pulls both grips inwards
    # On a free row F_i = -(r_i + reg U_i) with r = b - (K_ff + reg I) U_f (src/fea_solver.py:125), and
    # t_y.(K U) = (K t_y).U = 0, hence  R_top + R_bot - reg * sum_free U_y = sum_free r_y.
    free_y = torch.ones(K.n_rows, dtype=torch.bool, device=dev)
    free_y[kd_d] = False
    free_y[0::3] = False
    free_y[2::3] = False
    reg_term = fs.REGULARISATION * float(res.U[free_y].sum())
    bound = float(np.sqrt(int(free_y.sum()))) * tr * float(res.system.rhs.norm())
    gap = r_top + r_bot - reg_term
    print(f"2048^2 Y: it={res.iterations} true={tr:.2e} R_top={r_top:.6e} R_bot={r_bot:.6e} reg*sumU={reg_term:.2e} "
          f"|gap|={abs(gap):.2e} bound={bound:.2e} solve={res.ms_solve:.0f} ms")
    assert abs(gap) <= bound + 1e-10 * abs(r_top)
    assert abs(r_top + r_bot) <= 0.05 * abs(r_top)          # the 1e-12 I shift carries the rest (floating clusters)


def test_multigrid_and_block_jacobi_agree_at_full_size(big):
    """Above 512^2 the oracle's direct solve is out of reach, so parity is by the true residual (above) and by two
    independent solvers agreeing: the multigrid PCG (pcg_amg_kernel: V-cycle with FP32 level operators) and the
    block-Jacobi PCG (pcg_fused_kernel) share only the assembled K and the Dirichlet elimination.  Both to rtol 1e-12:
    same reaction force to 1e-8, same displacement field to 1e-6 (the system's conditioning times the tolerance)."""
    ctx, coords, n1, n2, mesh, K = big
    hi, lo = fs.grip_nodes(coords, 1.5, 1)
    kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, 1)
    a = fs.analyze_load_case(mesh, kd, kv, react_dofs=3 * hi + 1, rtol=1e-12, precond="amg", K=K)
    Ua, fa, ita = a.U.clone(), a.total_force, a.iterations
    assert a.system.precond == "amg"
    b = fs.analyze_load_case(mesh, kd, kv, react_dofs=3 * hi + 1, rtol=1e-12, precond="block6", K=K)
    assert b.system.precond == "block6" and b.iterations > 20 * ita
    assert abs(fa - b.total_force) <= 1e-8 * abs(b.total_force), (fa, b.total_force)
    err = float((Ua - b.U).norm() / b.U.norm())
    print(f"2048^2 Y: amg {ita} its, block6 {b.iterations} its, relL2(U) {err:.2e}, force {fa:.10e} / {b.total_force:.10e}")
    assert err <= 1e-6, err
