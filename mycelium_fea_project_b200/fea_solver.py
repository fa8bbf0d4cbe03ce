"""Drop-in for the reference's ``src/fea_solver.py`` / ``src/fea_solver_no_plotting.py``.

Same entry point (``fea_solver(results_dir, tol)``; CLI ``python -m
mycelium_fea_project_b200.fea_solver <results_dir>``), same constants, same
``results/sim_*`` input files and ``fea_results/`` output files, and the same three inner
functions a caller can use one at a time:

    bar_stiffness_bulk(p1s, p2s, E, A, I)            -> (K (N,6,6), L)       fea_solver.py:30
    assemble_global_stiffness(coords, elems, active) -> scipy.sparse.csr_matrix        :74
    solve_system(K, known_dofs, known_vals)          -> U                             :112

All arithmetic runs in hand-written sm_100a CUDA kernels through the C-ABI
(include/mycelium_fea.h); numpy / pandas / scipy appear only as the containers the reference's
signatures use.  The linear solve is a preconditioned CG (aggregation multigrid by default, or a
Jacobi / block-Jacobi variant) instead of SuperLU, run to ``PCG_RTOL`` (displacements agree with
the reference's direct solve to 1e-8 relative L2 -- tests/test_gpu_parity.py, tests/test_gpu_amg.py).
There is no CPU fallback.

The module-level constants are read at call time and may be overridden, because the
reference's committed goldens were produced with other values than the committed source
(SURVEY.md section 0.4).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from . import device as dv
from ._lib import MyceliumFeaError

# ---------------------------------------------------------------------------------------------
# Material & simulation parameters -- same expressions as src/fea_solver.py:14-28
# ---------------------------------------------------------------------------------------------
E_mod = 2500
d = 0.0002
t = 0.000001
A = 3.14 * ((d / 2) ** 2 - (d / 2 - t) ** 2)
I = A * 0.001
N_STEPS = 40
DISPLACEMENT_MAX = 0.02
MAX_STRAIN = 0.018
MAX_STRESS = E_mod * MAX_STRAIN
GRIP_LENGTH = 1.5
REGULARISATION = 1e-12          # fea_solver.py:125

# solver knobs (not in the reference, which uses a direct solve).  The ramp's failure switch
# |strain| > MAX_STRAIN is a hard threshold on the solution, so the drop-in's default tolerance is tighter than
# the 1e-10 the benchmark configurations name (bench.py passes its own rtol); with the multigrid
# preconditioner two more digits cost ~20 % more iterations.
PCG_RTOL = 1e-12
PCG_MAXIT = 500_000
# "amg" (aggregation multigrid, default; prepared as "block6" where the hierarchy is not applicable),
# "jacobi" (point), "block3" (3x3 node blocks), "block6" / "block12" (aligned blocks of 2 / 4 consecutive
# nodes).  MYC_PCG_PRECOND overrides the default.
PCG_PRECOND = os.environ.get("MYC_PCG_PRECOND", "amg")
if PCG_PRECOND not in ("amg", "jacobi", "block3", "block6", "block12"):
    raise ValueError(f"MYC_PCG_PRECOND={PCG_PRECOND!r}: expected amg, jacobi, block3, block6 or block12")


def _ctx():
    return dv.Context.get()


def _dev(a, dtype):
    arr = np.ascontiguousarray(a, dtype=dtype)
    return torch.from_numpy(arr if arr.flags.writeable else arr.copy()).to(_ctx().device)


# ---------------------------------------------------------------------------------------------
# the reference's three inner functions
# ---------------------------------------------------------------------------------------------
def bar_stiffness_bulk(p1s, p2s, E=None, A=None, I=None):
    """K_e (N,6,6) and L (N,) of N two-node bars; replaces fea_solver.py:30-68."""
    E = E_mod if E is None else E
    A_ = globals()["A"] if A is None else A
    I_ = globals()["I"] if I is None else I
    p1 = _dev(np.asarray(p1s, dtype=np.float64).reshape(-1, 3), np.float64)
    p2 = _dev(np.asarray(p2s, dtype=np.float64).reshape(-1, 3), np.float64)
    if p1.shape != p2.shape:
        raise ValueError("p1s and p2s must have the same shape")
    K, L = dv.bar_stiffness(_ctx(), p1, p2, E, A_, I_)
    return K.cpu().numpy(), L.cpu().numpy()


def _elem_columns(elems):
    """n1/n2 columns of the reference's elements DataFrame (or any mapping / (n1,n2) pair)."""
    if isinstance(elems, (tuple, list)) and len(elems) == 2:
        return np.asarray(elems[0]), np.asarray(elems[1])
    n1 = elems["n1"]
    n2 = elems["n2"]
    return np.asarray(getattr(n1, "values", n1)), np.asarray(getattr(n2, "values", n2))


def assemble_global_stiffness_device(coords, elems, active) -> dv.DeviceCSR:
    n1, n2 = _elem_columns(elems)
    mesh = dv.DeviceMesh.from_host(coords, n1, n2, active)
    return dv.assemble(_ctx(), mesh, E_mod, globals()["A"], globals()["I"])


def assemble_global_stiffness(coords, elems, active):
    """Global K as scipy CSR (int32 indices, sorted columns, duplicates summed, explicit zeros
    kept); replaces fea_solver.py:74-106."""
    return assemble_global_stiffness_device(coords, elems, active).to_scipy()


def solve_system(K, known_dofs, known_vals, return_info=False):
    """U (n_dof,) with U[known] = known_vals; replaces fea_solver.py:112-135.

    ``K`` may be a scipy sparse matrix (as the reference passes) or a DeviceCSR."""
    ctx = _ctx()
    Kd = K if isinstance(K, dv.DeviceCSR) else dv.DeviceCSR.from_scipy(K)
    kd = np.asarray(known_dofs, dtype=np.int64)
    if len(np.unique(kd)) != len(kd):
        raise ValueError("known_dofs contains duplicates")
    sysd = dv.apply_dirichlet(ctx, Kd, _dev(kd, np.int64), _dev(known_vals, np.float64), REGULARISATION,
                              precond=PCG_PRECOND)
    x, iters, relres = dv.pcg(ctx, Kd, sysd, precond=PCG_PRECOND, rtol=PCG_RTOL, maxit=PCG_MAXIT)
    U = dv.merge_solution(ctx, Kd, sysd, x).cpu().numpy()
    if return_info:
        return U, {"iterations": iters, "relres": relres, "true_relres": dv.true_residual(ctx, Kd, sysd, x)}
    return U


# ---------------------------------------------------------------------------------------------
# host logic of the step loop: grips and Dirichlet sets (fea_solver.py:205-210, 223-245)
# ---------------------------------------------------------------------------------------------
LOAD_CASES = {            # name -> (grip axis, prescribed component)
    "Y": (1, 1),          # the reference's only case: y grips, stretch in y
    "X": (0, 0),          # x grips, stretch in x
    "shear": (1, 0),      # y grips, prescribed x displacement
}


def grip_nodes(coords, tol=None, axis=1):
    """(hi, lo) node ids within ``tol`` of the max / min coordinate (fea_solver.py:205-210)."""
    tol = GRIP_LENGTH if tol is None else tol
    c = np.asarray(coords)[:, axis]
    ids = np.arange(len(c))
    return ids[np.abs(c - c.max()) < tol], ids[np.abs(c - c.min()) < tol]


def build_bc(hi_nodes, lo_nodes, d_hi, d_lo, comp=1):
    """known_dofs / known_vals with the reference's dict semantics (fea_solver.py:223-245):
    hi grip inserted first, then lo; a node in both keeps its first position and takes the lo
    value; the two non-prescribed components are clamped to 0."""
    hi = np.asarray(hi_nodes, dtype=np.int64)
    lo = np.asarray(lo_nodes, dtype=np.int64)
    lo_new = lo[~np.isin(lo, hi)]
    order = np.concatenate([hi, lo_new])
    val = np.where(np.isin(order, lo), float(d_lo), float(d_hi))
    known_dofs = (3 * order[:, None] + np.arange(3)).ravel()
    known_vals = np.zeros((len(order), 3))
    known_vals[:, comp] = val
    return known_dofs, known_vals.ravel()


# ---------------------------------------------------------------------------------------------
# device-resident load case (what the step loop and bench.py run)
# ---------------------------------------------------------------------------------------------
class LoadCaseResult:
    __slots__ = ("U", "total_force", "iterations", "relres", "K", "system", "x", "ms_assemble", "ms_setup", "ms_solve")


def analyze_load_case(mesh: dv.DeviceMesh, known_dofs, known_vals, react_dofs=None, x0=None,
                      rtol=None, precond=None, K=None, system=None) -> LoadCaseResult:
    """assemble -> Dirichlet (+ preconditioner setup) -> PCG -> U (-> reactions), everything staying on the
    device.  ``known_dofs``/``known_vals``/``react_dofs`` may be numpy arrays or device tensors.
    ``K`` / ``system``: reuse the matrix / the Dirichlet system (same known DOFs, new values; keeps its
    preconditioner) of an earlier call on the same mesh state -- the ramp's incremental path.
    Timings (CUDA events): ms_assemble, ms_setup (Dirichlet elimination + preconditioner), ms_solve (PCG,
    merge, reactions)."""
    ctx = _ctx()
    rtol = PCG_RTOL if rtol is None else rtol
    precond = PCG_PRECOND if precond is None else precond
    as_dev = lambda a, dt: a if isinstance(a, torch.Tensor) else _dev(a, dt)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    if K is None:
        K = dv.assemble(ctx, mesh, E_mod, globals()["A"], globals()["I"])
        system = None
    ev[1].record()
    sysd = dv.apply_dirichlet(ctx, K, as_dev(known_dofs, np.int64), as_dev(known_vals, np.float64),
                              REGULARISATION, precond=precond, reuse=system)
    ev[2].record()
    x, iters, relres = dv.pcg(ctx, K, sysd, x0=x0, precond=precond, rtol=rtol, maxit=PCG_MAXIT)
    out = LoadCaseResult()
    out.U = dv.merge_solution(ctx, K, sysd, x)
    out.total_force = None
    if react_dofs is not None:
        F = dv.spmv(ctx, K, out.U)                                       # fea_solver.py:257
        out.total_force = dv.gather_sum(ctx, F, as_dev(react_dofs, np.int64))   # :263-264
    ev[3].record()
    ev[3].synchronize()
    out.iterations, out.relres, out.K, out.system, out.x = iters, relres, K, sysd, x
    out.ms_assemble, out.ms_setup, out.ms_solve = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
    return out


# ---------------------------------------------------------------------------------------------
# the driver (fea_solver.py:186-335)
# ---------------------------------------------------------------------------------------------
def load_snapshot(results_dir):
    """nodes.csv / elements.csv of a results/sim_* directory (fea_solver.py:193-196) -> (coords, n1, n2).
    A ``mesh.npz`` side-car (snapshot_io / synth.write_snapshot) is preferred when it is at least as new as both
    CSVs (or when there are no CSVs); a side-car older than an edited CSV is ignored.
    The reference addresses grip nodes by the ``node_id`` column (:209-210) and elements by row position of the
    node table (:82-83); both agree only if node_id == row index, which is checked here."""
    import pandas as pd
    npz = os.path.join(results_dir, "mesh.npz")
    csvs = [os.path.join(results_dir, f) for f in ("nodes.csv", "elements.csv")]
    have_csv = all(os.path.isfile(c) for c in csvs)
    if os.path.isfile(npz) and (not have_csv or os.path.getmtime(npz) >= max(os.path.getmtime(c) for c in csvs)):
        m = np.load(npz)
        return m["coords"], m["n1"], m["n2"]
    nodes = pd.read_csv(csvs[0])
    elems = pd.read_csv(csvs[1])
    if "node_id" in nodes.columns and not np.array_equal(nodes["node_id"].values, np.arange(len(nodes))):
        raise ValueError(f"{csvs[0]}: node_id must equal the row index (the reference mixes both addressings)")
    return nodes[["x", "y", "z"]].values, elems["n1"].values, elems["n2"].values


def fea_ramp(coords, n1, n2, tol=None, load_case="Y", warm_start=True, verbose=False, incremental=True):
    """The displacement ramp on in-memory arrays; returns the per-step records
    (stress, active, disp, force_disp) the reference accumulates (fea_solver.py:200-203).
    ``incremental``: reuse K and the preconditioner while the set of active elements is unchanged."""
    ctx = _ctx()
    tol = GRIP_LENGTH if tol is None else tol
    axis, comp = LOAD_CASES[load_case]
    coords = np.asarray(coords, dtype=np.float64)
    mesh = dv.DeviceMesh.from_host(coords, n1, n2)
    hi, lo = grip_nodes(coords, tol, axis)
    react = _dev(3 * hi + comp, np.int64)
    rec = {"stress": [], "active": [], "disp": [], "force_disp": [], "iterations": [], "reassembled": [],
           "solve_seconds": []}
    x_prev, step_prev, topology_changed = None, 0, False
    K_prev = sys_prev = None
    for step in range(N_STEPS):
        f = step / (N_STEPS - 1)
        d_hi, d_lo = +DISPLACEMENT_MAX * f, -DISPLACEMENT_MAX * f
        if verbose:
            print(f"Step {step + 1}/{N_STEPS} | d_hi={d_hi:.3f}, d_lo={d_lo:.3f}")
        known_dofs, known_vals = build_bc(hi, lo, d_hi, d_lo, comp)
        x0 = None
        if warm_start and x_prev is not None and step_prev > 0 and not topology_changed:
            # The response is linear in the load factor while no element fails, so the scaled
            # previous solution is already converged.  After a failure we restart from 0: a
            # warm start would leave stale displacements on pieces that have become detached
            # (their only stiffness is the 1e-12 shift, invisible to the residual), whereas
            # the reference's direct solve returns exactly 0 there.
            x0 = x_prev * (step / step_prev)
        # Incremental path: while no element fails, K, its Dirichlet structure and the preconditioner (multigrid
        # hierarchy / block inverses) are unchanged -- only the prescribed values move -- so neither the assembly
        # nor the preconditioner setup is redone (the reference re-assembles every step, :220).
        reuse = incremental and K_prev is not None and not topology_changed
        t_solve = time.time()
        try:
            res = analyze_load_case(mesh, known_dofs, known_vals, react_dofs=react, x0=x0,
                                    K=K_prev if reuse else None, system=sys_prev if reuse else None)
        except MyceliumFeaError as exc:              # the reference's LinAlgError branch (:250-254)
            print(f"Solver failure at step {step + 1}: {exc}. Saving partial results and stopping.")
            break
        # what the reference's solve_runtime.txt times (:247-261: solve_system + K @ U); here it also covers the
        # (re-)assembly and preconditioner setup of the step, which run inside the same call
        rec["solve_seconds"].append(time.time() - t_solve)
        rec["reassembled"].append(not reuse)
        K_prev, sys_prev = res.K, res.system
        x_prev, step_prev = res.x, step
        rec["force_disp"].append([d_hi - d_lo, res.total_force])
        n_before = int(mesh.active.sum().item())
        stress, n_active = dv.strain_update(ctx, mesh, res.U, E_mod, MAX_STRAIN)     # :269-284
        topology_changed = n_active != n_before
        rec["stress"].append(stress.cpu().numpy())
        rec["active"].append(mesh.active.cpu().numpy().astype(bool))
        rec["disp"].append(res.U.cpu().numpy())
        rec["iterations"].append(res.iterations)
        if n_active == 0:
            if verbose:
                print(f"Simulation stopped early at step {step + 1}.")
            break
    return rec


def fea_ramp_distributed(coords, n1, n2, tol=None, load_case="Y", warm_start=True, verbose=False, incremental=True):
    """The displacement ramp on N GPUs (one process per GPU, torch.distributed initialised with the nccl backend;
    collective).  Same loop, same shortcuts and same records as ``fea_ramp``; every rank assembles and solves its row
    block (dist.DistributedSolver), evaluates strain / failure on its own elements -- an element cut by the partition
    lives on both sides, which take the same decision from the same gathered U -- and the per-step records are
    combined on every rank (the caller normally writes them on rank 0 only)."""
    import torch.distributed as tdist
    from . import dist as md
    ctx = _ctx()
    tol = GRIP_LENGTH if tol is None else tol
    axis, comp = LOAD_CASES[load_case]
    coords = np.asarray(coords, dtype=np.float64)
    n_elem = len(n1)
    solver = md.DistributedSolver((coords, n1, n2), device=ctx.device)
    mesh = solver.mesh
    eidx = torch.from_numpy(solver.elem_index).to(ctx.device)
    hi, lo = grip_nodes(coords, tol, axis)
    react = 3 * hi + comp
    rec = {"stress": [], "active": [], "disp": [], "force_disp": [], "iterations": [], "reassembled": [],
           "solve_seconds": []}
    x_prev, step_prev, topology_changed = None, 0, False
    K_prev = sys_prev = None
    for step in range(N_STEPS):
        f = step / (N_STEPS - 1)
        d_hi, d_lo = +DISPLACEMENT_MAX * f, -DISPLACEMENT_MAX * f
        if verbose and solver.rank == 0:
            print(f"Step {step + 1}/{N_STEPS} | d_hi={d_hi:.3f}, d_lo={d_lo:.3f}")
        known_dofs, known_vals = build_bc(hi, lo, d_hi, d_lo, comp)
        x0 = None
        if warm_start and x_prev is not None and step_prev > 0 and not topology_changed:
            x0 = x_prev * (step / step_prev)                 # see fea_ramp
        reuse = incremental and K_prev is not None and not topology_changed
        t_solve = time.time()
        K = K_prev if reuse else solver.assemble(E_mod, globals()["A"], globals()["I"])
        try:
            out = solver.load_case(K, known_dofs, known_vals, react_dofs=react, rtol=PCG_RTOL, precond=PCG_PRECOND,
                                   maxit=PCG_MAXIT, reg=REGULARISATION, gather_U=True, system=sys_prev if reuse else None, x0=x0)
        except MyceliumFeaError as exc:
            if solver.rank == 0:
                print(f"Solver failure at step {step + 1}: {exc}. Saving partial results and stopping.")
            break
        rec["solve_seconds"].append(time.time() - t_solve)
        rec["reassembled"].append(not reuse)
        K_prev, sys_prev = K, out["system"]
        x_prev, step_prev = out["x"], step
        rec["force_disp"].append([d_hi - d_lo, out["total_force"]])
        n_before = int(mesh.active.sum().item())
        stress, n_active = dv.strain_update(ctx, mesh, out["U"], E_mod, MAX_STRAIN)       # this rank's elements
        # combine: every element lives on at least one rank, cut elements on two with identical values
        stress_g = torch.full((n_elem,), -float("inf"), dtype=torch.float64, device=ctx.device)
        stress_g[eidx] = stress
        active_g = torch.zeros((n_elem,), dtype=torch.int32, device=ctx.device)
        active_g[eidx] = mesh.active.to(torch.int32)
        flags = torch.tensor([1 if n_active != n_before else 0, n_active], dtype=torch.int64, device=ctx.device)
        tdist.all_reduce(stress_g, op=tdist.ReduceOp.MAX)
        tdist.all_reduce(active_g, op=tdist.ReduceOp.MAX)
        tdist.all_reduce(flags, op=tdist.ReduceOp.SUM)
        topology_changed = int(flags[0].item()) > 0
        rec["stress"].append(stress_g.cpu().numpy())
        rec["active"].append(active_g.cpu().numpy().astype(bool))
        rec["disp"].append(out["U"].cpu().numpy())
        rec["iterations"].append(out["iterations"])
        if int(flags[1].item()) == 0:
            if verbose and solver.rank == 0:
                print(f"Simulation stopped early at step {step + 1}.")
            break
    return rec


def _write_wide_csv(path, header, rows, steps):
    """One CSV line per load step: the values of ``rows[k]`` followed by the 1-based step number -- byte for byte what
    ``pandas.DataFrame(rows, columns=header[:-1]).assign(step=steps).to_csv(path, index=False)`` writes (floats in
    their shortest round-trip form, booleans as True / False), without building a 22,125-column DataFrame: at the
    reference's sizes the CSV formatting, not the solve, was the ramp's wall clock (CPU test: test_host_io.py).
    Returns False (nothing written) if the data needs pandas' special cases (no rows, non-finite values)."""
    if len(rows) == 0:
        return False
    a = np.asarray(rows)
    if a.ndim != 2 or a.shape[1] != len(header) - 1:
        return False
    if a.dtype == np.bool_:
        lines = [",".join(r) for r in np.where(a, "True", "False").tolist()]
    elif a.dtype == np.float64:
        if not np.isfinite(a).all():
            return False                       # pandas writes NaN as an empty field
        lines = [",".join(map(repr, r)) for r in a.tolist()]
    else:
        return False
    with open(path, "w", newline="") as f:
        f.write(",".join(header) + "\n")
        f.write("".join(f"{ln},{int(st)}\n" for ln, st in zip(lines, steps)))
    return True


def write_results(fea_dir, rec, n_elems, total_time=None):
    """The reference's four CSVs + runtime.txt (fea_solver.py:298-333), same columns, same bytes."""
    import pandas as pd
    os.makedirs(fea_dir, exist_ok=True)
    cols = [f"elem_{i}" for i in range(n_elems)]
    steps = np.arange(1, len(rec["stress"]) + 1)
    n_dof = len(rec["disp"][0]) if rec["disp"] else 0
    for name, rows, header in (("stress_record.csv", rec["stress"], cols),
                               ("active_elements.csv", rec["active"], cols),
                               ("node_displacements.csv", rec["disp"], [str(i) for i in range(n_dof)])):
        path = os.path.join(fea_dir, name)
        if not _write_wide_csv(path, header + ["step"], rows, steps):
            df = pd.DataFrame(rows, columns=cols if header is cols else np.arange(n_dof))
            df["step"] = steps
            df.to_csv(path, index=False)
    pd.DataFrame(rec["force_disp"], columns=["total_displacement", "total_force"]).to_csv(
        os.path.join(fea_dir, "force_displacement.csv"), index=False)
    if total_time is not None:
        with open(os.path.join(fea_dir, "runtime.txt"), "w") as f:
            f.write(f"Total FEA runtime: {total_time:.6f} seconds\n")


def fea_solver(results_dir, tol=None, load_case="Y", binary_outputs=None):
    """Run the 40-step ramp on ``results_dir`` and write ``fea_results/``; same contract as
    the reference's fea_solver(results_dir, tol=GRIP_LENGTH) (fea_solver.py:186)."""
    start = time.time()
    print(f"Running FEA on geometry from {results_dir}")
    fea_dir = os.path.join(results_dir, "fea_results")
    os.makedirs(fea_dir, exist_ok=True)
    coords, n1, n2 = load_snapshot(results_dir)
    import torch.distributed as tdist
    multi = tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1
    if multi:                                     # launched with torch.distributed.run: row-partitioned over the ranks
        rec = fea_ramp_distributed(coords, n1, n2, tol=tol, load_case=load_case, verbose=True)
        if tdist.get_rank() != 0:
            return rec                            # rank 0 writes the files
    else:
        rec = fea_ramp(coords, n1, n2, tol=tol, load_case=load_case, verbose=True)
    n_dof = 3 * len(coords)
    if binary_outputs is None:
        binary_outputs = n_dof > 2_000_000       # one CSV column per DOF is unusable beyond this
    if binary_outputs:
        np.savez_compressed(os.path.join(fea_dir, "records.npz"), stress=np.array(rec["stress"]),
                            active=np.array(rec["active"]), disp=np.array(rec["disp"]),
                            force_disp=np.array(rec["force_disp"]))
    else:
        write_results(fea_dir, rec, len(n1))
    with open(os.path.join(fea_dir, "solve_runtime.txt"), "w") as f:          # fea_solver.py:214-215, 260-261
        f.write("step, runtime_s\n")
        for k, t in enumerate(rec["solve_seconds"]):
            f.write(f"{k + 1}, {t:.6f}\n")
    total = time.time() - start
    with open(os.path.join(fea_dir, "runtime.txt"), "w") as f:
        f.write(f"Total FEA runtime: {total:.6f} seconds\n")
    print(f"FEA completed. Results saved to {fea_dir}")
    print(f"Total runtime: {total:.3f} seconds")
    return rec


if __name__ == "__main__":
    if len(sys.argv) < 2:
        print("Usage: python fea_solver.py <results_dir>      (N GPUs: python -m torch.distributed.run "
              "--nproc-per-node N -m mycelium_fea_project_b200.fea_solver <results_dir>)")
        sys.exit()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:          # launched by torch.distributed.run: one rank per GPU
        import torch.distributed as tdist
        from . import dist as md
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        try:
            fea_solver(sys.argv[1], tol=GRIP_LENGTH)
            md.shutdown()
        finally:
            tdist.destroy_process_group()
    else:
        fea_solver(sys.argv[1], tol=GRIP_LENGTH)
