#!/usr/bin/env python
"""Benchmark of the FEA hot path (BASELINE.json metric: assemble + PCG solve time and MDOF/s at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid 2048] [--precond amg]

Workload (config.workload): BASELINE.json configs[2] -- the synthetic 2048 x 2048 mycelium occupancy grid
(8.39 M DOF, operator 1.3 GB: 10x the 126 MB L2, so nothing in a step is cache-resident), X and Y load cases.
One "step" = both load cases, each one a full pass of the hot path: assemble K from the mesh (element stiffness
-> CSR), Dirichlet elimination, preconditioner setup, PCG to rtol 1e-10, reaction sum.
value = DOFs solved per second = (load cases x n_dof) / step time.  ms_assemble / ms_setup / ms_solve split a step
(CUDA events, summed over its load cases).
With N GPUs the specimen's cross-section is N times larger at the same gauge length (Y case: 2048 rows x 2048 N
columns; X case: 2048 N rows x 2048 columns), row-partitioned over the ranks, so per-GPU work is fixed ("weak").

Timing: W (>= 3) warm-up steps, then K steps between barrier + cuda synchronize, CUDA events, max over ranks.
`value` starts with the mesh and BCs resident in HBM; `e2e` runs the same step through the host-buffer entry
point (N = 1: the C-ABI call myc_load_case_host on pinned host arrays; N > 1: DistributedSolver on pinned host
arrays), H2D of mesh + BCs and D2H of U inside the timed region.

Extra records in the same JSON line (each one-shot: 1 warm + 1 timed run, not part of `value`):
  strong_4096      BASELINE configs[3]: the SAME 4096^2 mesh (33.5 M DOF), X load case, solved on all N GPUs --
                   strong scaling across the --gpus runs; at N > 1 with `parity`: the same solve repeated on ONE
                   GPU (rank 0, private context) and compared (relative L2 of U, reaction force)
  config4_8192     (N = 8 only) BASELINE configs[4]: 8192^2 (134 M DOF), Y and shear load cases
  roofline_block6  (N = 1) the 2048^2 Y load case with the block-Jacobi PCG (pcg_fused_kernel), the kernel the
                   previous round's roofline was quoted on, now on an operator that cannot sit in L2
  roofline_hbm     (N = 1) the CSR SpMV alone on the 2048^2 operator, L2 flushed between launches
  ramp             (N = 1) the reference's own published workload: the 40-step displacement ramp of
                   results/sim_20251117_181147 through fea_solver() end to end (CSV in, CSVs out)
  cpu_baseline     (N = 1) the reference's CPU path on the box's host cores, bounded samples

--impl reference times the reference's CPU path (oracle/: scipy COO->CSR assembly + SuperLU spsolve, the
restatement of src/fea_solver.py pinned to the reference's goldens) on the host cores: the same X + Y step on the
2048^2 specimen, every element through the reference's literal assembly loop, nothing extrapolated; `steps` is
the number of steps actually run.
"""
import argparse
import ctypes as C
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRIP = 1.5
DISP = 0.02
RTOL = 1e-10
RTOL_ONE_SHOT = 1e-12   # the strong-scaling / configs[4] records are compared ACROSS runs (1 vs N GPUs): two more digits
MAXIT = 400_000      # bound on PCG iterations (a mis-set problem must not burn GPU minutes)
NOMINAL_HBM_GBS = 8000.0   # the north_star's "~8 TB/s": fractions are quoted against the measured copy peak AND this
PUBLISHED_RAMP_S = 71.76   # /root/reference/results/sim_20251117_181147/fea_results/runtime.txt:1 (incl. plotting)


def specimen(case, grid, n_gpus, seed=0):
    from mycelium_fea_project_b200.synth import synth_network
    if case == "X":
        return synth_network(grid * n_gpus, grid, seed=seed)
    return synth_network(grid, grid * n_gpus, seed=seed)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key, iterations):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture -- quoted only if the
    capture is of the same solve (same iteration count), otherwise stale and therefore null."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.isfile(p):
        return None, "no committed capture"
    rec = json.load(open(p)).get(key)
    if not rec:
        return None, f"no committed capture for {key}"
    if rec.get("iterations") != iterations:
        return None, f"committed capture ({rec.get('source')}) is of a {rec.get('iterations')}-iteration launch, this run took {iterations}"
    return rec["dram_bytes_per_launch"], rec.get("source")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def workload_config(args, n_dof):
    return {"workload": f"synthetic {args.grid}x{args.grid} mycelium occupancy grid per GPU (BASELINE configs[2]), "
                        f"{' and '.join(n_dof.keys())} load cases, specimen cross-section x{args.gpus}",
            "grid": args.grid, "n_dof": n_dof, "load_cases": list(n_dof.keys()),
            "solver": f"{args.precond}-PCG", "rtol": RTOL, "grip_length": GRIP, "seed": 0,
            "l2": "inputs larger than L2: the operator of one load case is 1.3 GB per GPU (126 MB L2), no flush needed",
            "parallelism": f"row-partition x{args.gpus}"}


# =================================================================================================
# reference arm / CPU baselines (the only place that executes oracle/)
# =================================================================================================
def _reference_load_case(fo, mesh, case, literal_elems=None):
    """One load case of the reference's CPU path.  Returns seconds: (literal assembly loop, vectorised bit-identical
    assembly, everything after the assembly).  literal_elems: time the literal loop on that many elements only
    (None: all of them)."""
    coords, n1, n2 = mesh
    axis, comp = {"Y": (1, 1), "X": (0, 0)}[case]
    active = np.ones(len(n1), bool)
    t0 = time.perf_counter()
    K = fo.assemble_global_stiffness(coords, n1, n2, active)
    t_vec = time.perf_counter() - t0
    if literal_elems is None or literal_elems >= len(n1):
        t0 = time.perf_counter()
        fo.assemble_global_stiffness_loop(coords, n1, n2, active)       # src/fea_solver.py:93-105, every element
        t_loop, loop_elems = time.perf_counter() - t0, len(n1)
    else:
        sub = np.zeros(len(n1), bool)
        sub[:literal_elems] = True
        t0 = time.perf_counter()
        fo.assemble_global_stiffness_loop(coords, n1, n2, sub)
        t_loop, loop_elems = time.perf_counter() - t0, literal_elems
    t0 = time.perf_counter()
    hi, lo = fo.grip_nodes(coords, GRIP, axis)
    kd, kv = fo.build_bc(hi, lo, DISP, -DISP, comp)
    U = fo.solve_system(K, kd, kv)
    _ = (K @ U)[3 * hi + comp].sum()
    t_rest = time.perf_counter() - t0
    return t_loop, loop_elems, t_vec, t_rest


def host_probe():
    """Is the reference's PETSc path runnable on this box? (north_star: 'both timed on the box's own host cores')"""
    return {"mpirun": shutil.which("mpirun"), "mpicxx": shutil.which("mpicxx"), "PETSC_DIR": os.environ.get("PETSC_DIR"),
            "petsc_pkg_config": subprocess.run(["sh", "-c", "pkg-config --exists PETSc petsc 2>/dev/null && echo yes || echo no"],
                                               capture_output=True, text=True).stdout.strip(),
            "verdict": "PETSc / MPI unavailable on this box: the PETSc variant is represented by oracle/pcg_port.c "
                       "(C + OpenMP restatement of MatZeroRowsColumns + KSPCG/PCJACOBI)"
            if not (shutil.which("mpirun") and os.environ.get("PETSC_DIR")) else "PETSc toolchain present"}


def run_reference(args):
    """The reference's CPU path on the host cores (rank 0 only): literal 36-append assembly loop over every
    element + scipy COO->CSR + Dirichlet reduction + SuperLU spsolve + K@U, the oracle port of src/fea_solver.py."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fea_oracle as fo
    cases = ["X", "Y"]
    # the per-GPU share of the weak-scaling specimen (= the whole specimen at N = 1): the CPU path's throughput
    # only falls with size (super-linear direct solve), so this does not understate the reference
    meshes = {c: specimen(c, args.grid, 1) for c in cases}
    n_dof_sample = {c: 3 * len(meshes[c][0]) for c in cases}
    n_dof_cfg = {c: 3 * len(meshes[c][0]) for c in cases} if args.gpus == 1 else \
        {c: None for c in cases}
    if args.gpus > 1:            # the config must name the same specimen as the GPU arm's line
        for c in cases:
            n_dof_cfg[c] = 3 * len(specimen(c, args.grid, args.gpus)[0])
    budget_s = args.reference_budget
    t_start = time.perf_counter()
    steps_run, t_lit, t_res = 0, 0.0, 0.0
    for _ in range(max(1, args.steps)):
        a = b = 0.0
        for c in cases:
            t_loop, _, t_vec, t_rest = _reference_load_case(fo, meshes[c], c)
            a += t_loop + t_rest                 # the literal loop's function ends in the same csr_matrix(...) call
            b += t_vec + t_rest
        t_lit += a
        t_res += b
        steps_run += 1
        if (time.perf_counter() - t_start) * (steps_run + 1) / steps_run > budget_s:
            break
    t_lit /= steps_run
    t_res /= steps_run
    total = sum(n_dof_sample.values())
    value = total / t_lit / 1e6
    sample = (f"{steps_run} step(s) of the X + Y load cases on the {args.grid}x{args.grid} specimen"
              + ("" if args.gpus == 1 else f" (one GPU's share of the x{args.gpus} specimen)")
              + ": every element through the reference's literal 36-append assembly loop, scipy COO->CSR, Dirichlet "
                "reduction, SuperLU spsolve and K@U, single-threaded as in the reference; nothing extrapolated. "
                "restated_value: same with the vectorised, bit-identical assembly")
    line = {
        "impl": "reference", "metric": "assemble+solve throughput", "value": value, "unit": "MDOF/s",
        "n_gpus": args.gpus, "steps": steps_run, "warmup": 0, "steps_requested": args.steps,
        "ms_per_step": t_lit * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_dof_cfg),
        "cpu_baseline": {"value": value, "unit": "MDOF/s", "cores": 1, "kind": "port", "sample": sample,
                         "restated_value": total / t_res / 1e6, "restated_ms_per_step": t_res * 1e3,
                         "host_cores_available": os.cpu_count(), "host_probe": host_probe()},
        "e2e": {"value": value, "unit": "MDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def petsc_style_sample(fo, mesh, n_iters=100):
    """Bounded sample of the reference's PETSc path (MatZeroRowsColumns + KSPCG/PCJACOBI, src/fea_petsc.cpp:303-341)
    restated in C/OpenMP (oracle/pcg_port.c; PETSc itself is not in this image): n_iters iterations on all host
    threads.  Reported per iteration -- no extrapolation to a full solve."""
    from oracle import pcg_port
    if not pcg_port.available():
        return {"unavailable": "oracle/_build/libpcg_port.so not built"}
    coords, n1, n2 = mesh
    K = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
    kd, kv = fo.build_bc(*fo.grip_nodes(coords, GRIP, 1), DISP, -DISP, 1)
    pcg_port.solve_system_petsc_style(K, kd, kv, rtol=RTOL, max_iters=10)          # touch memory
    t0 = time.perf_counter()
    _, it, _ = pcg_port.solve_system_petsc_style(K, kd, kv, rtol=RTOL, max_iters=n_iters)
    dt = time.perf_counter() - t0
    return {"kind": "port", "threads": pcg_port.threads(), "n_dof": K.shape[0], "ms_per_iteration": dt / max(it, 1) * 1e3,
            "iterations_timed": it,
            "sample": f"{n_iters} point-Jacobi PCG iterations on the Y operator with all host threads (timing "
                      "includes the Dirichlet elimination pass); a full solve needs tens of thousands of them"}


def cpu_baseline(args):
    """Bounded CPU samples of the same workload on the box's host (rank 0, N = 1): ~20-30 s in total."""
    from oracle import fea_oracle as fo
    g = max(64, args.grid // 2)
    mesh = specimen("Y", g, 1)
    n_dof = 3 * len(mesh[0])
    lit = 100_000
    t_loop, loop_elems, t_vec, t_rest = _reference_load_case(fo, mesh, "Y", literal_elems=lit)
    per_elem = t_loop / loop_elems
    t_ref = per_elem * len(mesh[1]) + t_rest
    out = {"value": n_dof / t_ref / 1e6, "unit": "MDOF/s", "cores": 1, "kind": "port",
           "restated_value": n_dof / (t_vec + t_rest) / 1e6,
           "seconds": {"literal_loop_us_per_element": per_elem * 1e6, "vectorised_assembly": t_vec, "dirichlet_spsolve_reactions": t_rest},
           "sample": f"Y load case on the {g}x{g} grid ({n_dof} DOF, a quarter of the benchmark specimen -- the direct "
                     f"solve is super-linear, so the full size is slower per DOF): literal 36-append assembly loop timed "
                     f"on {loop_elems} of {len(mesh[1])} elements and scaled (it is element-by-element), scipy COO->CSR, "
                     "Dirichlet reduction, SuperLU spsolve, K@U -- single-threaded as in the reference. The full-size, "
                     "nothing-extrapolated measurement is `bench.py --impl reference`",
           "host_cores_available": os.cpu_count(), "host_probe": host_probe()}
    try:
        out["petsc_style_pcg"] = petsc_style_sample(fo, mesh)
    except Exception as exc:                       # the baseline must never break the bench line
        out["petsc_style_pcg"] = {"unavailable": repr(exc)}
    return out


def _unpack_real_snapshot(dst):
    src = os.path.join(ROOT, "tests", "golden", "ref_results", "sim_20251117_181147")
    os.makedirs(dst, exist_ok=True)
    for f in ("nodes.csv", "elements.csv"):
        with gzip.open(os.path.join(src, f + ".gz"), "rb") as fi, open(os.path.join(dst, f), "wb") as fo_:
            shutil.copyfileobj(fi, fo_)
    return src


def ramp_record(fs, cpu=True):
    """The reference's published workload: fea_solver(results_dir) on results/sim_20251117_181147 (7,375 nodes,
    22,125 DOF, 40 steps with the failure cascade), CSV in -> four CSVs out, wall clock as runtime.txt reports it."""
    import contextlib
    import io
    import pandas as pd
    out = {"workload": "results/sim_20251117_181147, committed constants, 40-step ramp, fea_solver() end to end",
           "published_reference_seconds": PUBLISHED_RAMP_S,
           "published_source": "reference results/sim_20251117_181147/fea_results/runtime.txt:1 (with PNG plotting; "
                               "50.8 s without plotting measured on the build container, BASELINE.md section 2)"}
    with tempfile.TemporaryDirectory() as tmp:
        src = _unpack_real_snapshot(tmp)
        with contextlib.redirect_stdout(io.StringIO()):
            fs.fea_solver(tmp, tol=fs.GRIP_LENGTH)                       # warm-up (allocations, first launches)
            t0 = time.perf_counter()
            rec = fs.fea_solver(tmp, tol=fs.GRIP_LENGTH)
            out["seconds"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            coords, n1, n2 = fs.load_snapshot(tmp)
            rec2 = fs.fea_ramp(coords, n1, n2)
            out["seconds_compute_only"] = time.perf_counter() - t0
        gold = pd.read_csv(os.path.join(src, "fea_results", "force_displacement.csv"), float_precision="round_trip").values
        got = np.array(rec["force_disp"])
        out["steps"] = len(rec["stress"])
        out["force_curve_max_rel_diff_vs_reference_csv"] = float(np.abs(got[:, 1] - gold[:, 1]).max() / np.abs(gold[:, 1]).max())
        out["iterations_per_step"] = [int(i) for i in rec2["iterations"]]
        out["reassembled_steps"] = int(sum(rec2["reassembled"]))
        if cpu:
            from oracle import fea_oracle as fo
            t0 = time.perf_counter()
            res = fo.fea_ramp(coords, n1, n2)
            out["cpu_oracle_seconds"] = time.perf_counter() - t0
            out["cpu_oracle_sample"] = ("all 40 steps of the oracle ramp on one host core, in memory (no CSV I/O): VECTORISED "
                                        "bit-identical assembly + spsolve + strain update -- the reference's literal assembly "
                                        "loop alone adds ~0.2 s per step (BASELINE.md: 9.1 s of 37 s)")
            out["cpu_oracle_cascade_equal"] = bool(np.array_equal(np.array(res.active), np.array(rec2["active"])))
    return out


# =================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mycelium_fea_project_b200 import device as dv, fea_solver as fs, dist as md
    from mycelium_fea_project_b200._lib import lib, check, PRECONDITIONERS

    ctx = dv.Context.get(torch.device("cuda", local))
    dev = ctx.device
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def make_problem(coords, n1, n2, case):
        axis, comp = fs.LOAD_CASES[case]
        hi, lo = fs.grip_nodes(coords, GRIP, axis)
        kd, kv = fs.build_bc(hi, lo, DISP, -DISP, comp)
        react = (3 * hi + comp).astype(np.int64)
        p = {"coords": coords, "n1": n1, "n2": n2, "kd": kd, "kv": kv, "react": react, "n_dof": 3 * len(coords), "case": case}
        if world > 1:
            p["solver"] = md.DistributedSolver((coords, n1, n2), device=dev)
            p["mesh"] = p["solver"].mesh
        else:
            p["mesh"] = dv.DeviceMesh.from_host(coords, n1, n2)
        p["kd_d"] = torch.from_numpy(kd).to(dev)
        p["kv_d"] = torch.from_numpy(kv).to(dev)
        p["react_d"] = torch.from_numpy(react).to(dev)
        return p

    def solve_problem(p, gather_U=False, rtol=RTOL):
        """One pass of the hot path on device-resident inputs.  Returns a dict of results + timings."""
        if world > 1:
            s = p["solver"]
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            K = s.assemble(fs.E_mod, fs.A, fs.I)
            ev[1].record()
            r = s.load_case(K, p["kd_d"], p["kv_d"], react_dofs=p["react"], rtol=rtol, precond=args.precond,
                            gather_U=gather_U, maxit=MAXIT)
            out = {"iterations": r["iterations"], "relres": r["relres"], "total_force": r["total_force"],
                   "ms_assemble": ev[0].elapsed_time(ev[1]), "ms_setup": r["ms_setup"], "ms_solve": r["ms_solve"],
                   "precond": r["precond"], "nnz_local": K.nnz}
            p["last"] = (K, r)
        else:
            r = fs.analyze_load_case(p["mesh"], p["kd_d"], p["kv_d"], react_dofs=p["react_d"], rtol=rtol,
                                     precond=args.precond)
            out = {"iterations": r.iterations, "relres": r.relres, "total_force": r.total_force,
                   "ms_assemble": r.ms_assemble, "ms_setup": r.ms_setup, "ms_solve": r.ms_solve,
                   "precond": r.system.precond, "nnz": r.K.nnz}
            p["last"] = r
        return out

    def true_relres(p):
        if world > 1:
            K, r = p["last"]
            return p["solver"].true_residual(K, r["system"], r["x"])
        r = p["last"]
        return dv.true_residual(ctx, r.K, r.system, r.x)

    # ---------------------------------------------------------------------------------------------
    cases = ["X", "Y"]
    prob = {c: make_problem(*specimen(c, args.grid, world), c) for c in cases}
    n_dof = {c: prob[c]["n_dof"] for c in cases}
    info = {}
    split = {"ms_assemble": 0.0, "ms_setup": 0.0, "ms_solve": 0.0}
    counting = [False]

    trace = os.environ.get("BENCH_TRACE") == "1"

    def device_step():
        for c in cases:
            info[c] = solve_problem(prob[c])
            if trace and rank == 0:
                print(f"[trace] {c} assemble {info[c]['ms_assemble']:.1f} setup {info[c]['ms_setup']:.1f} solve "
                      f"{info[c]['ms_solve']:.1f} its {info[c]['iterations']}", file=sys.stderr, flush=True)
            if counting[0]:
                for k in split:
                    split[k] += info[c][k]

    for _ in range(warmup):
        device_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.myc_profile_reset(ctx.h, 1)
    launches0 = ctx.launches
    counting[0] = True
    ms_total = timed(device_step, args.steps)
    counting[0] = False
    launches = ctx.launches - launches0
    prof = (C.c_double * 4)()
    lib.myc_profile_get(ctx.h, prof)
    lib.myc_profile_reset(ctx.h, 0)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    total_dof = sum(n_dof.values())
    value = total_dof / (ms_step * 1e-3) / 1e6
    split = {k: max_over_ranks(v / args.steps) for k, v in split.items()}
    for c in cases:
        info[c]["true_relres"] = true_relres(prob[c])
        if world == 1 and info[c]["precond"] == "amg":
            info[c]["amg_levels"] = [list(l) for l in dv.amg_levels(ctx)[0]] if c == cases[-1] else None

    # ---- e2e: host buffers in, U out, every step
    h2d = d2h = 0
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    if world == 1:
        host = {}
        for c in cases:
            p = prob[c]
            host[c] = {"coords": pin(p["coords"]), "n1": pin(p["n1"].astype(np.int32)), "n2": pin(p["n2"].astype(np.int32)),
                       "kd": pin(p["kd"]), "kv": pin(p["kv"]), "react": pin(p["react"]),
                       "U": torch.empty(p["n_dof"], dtype=torch.float64).pin_memory()}
            h2d += sum(host[c][k].numel() * host[c][k].element_size() for k in ("coords", "n1", "n2", "kd", "kv", "react"))
            d2h += host[c]["U"].numel() * 8 + 8
        ptr = lambda t: C.c_void_p(t.data_ptr())
        pc = PRECONDITIONERS[args.precond]

        def e2e_step():
            for c in cases:
                h = host[c]
                force, iters, rel, nnz = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
                check(ctx.h, lib.myc_load_case_host(
                    ctx.h, ptr(h["coords"]), ptr(h["n1"]), ptr(h["n2"]), None, h["n1"].numel(), h["coords"].shape[0],
                    float(fs.E_mod), fs.A, fs.I, ptr(h["kd"]), ptr(h["kv"]), h["kd"].numel(), 1e-12, pc, RTOL, MAXIT,
                    ptr(h["react"]), h["react"].numel(), ptr(h["U"]), C.byref(force), C.byref(iters), C.byref(rel),
                    C.byref(nnz), None, None))
                info[c]["e2e_total_force"] = force.value
        api = "myc_load_case_host (C-ABI, pinned host buffers)"
    else:
        pinned = {}
        for c in cases:
            p = prob[c]
            s = p["solver"]
            own = 3 * (s.plan.node_end - s.plan.node_begin)
            pinned[c] = {"coords": pin(np.asarray(p["coords"], dtype=np.float64)[s.coord_lo:s.coord_hi]),
                         "n1": pin(np.asarray(p["n1"])[s.elem_index].astype(np.int32)),
                         "n2": pin(np.asarray(p["n2"])[s.elem_index].astype(np.int32)),
                         "kd": pin(p["kd"]), "kv": pin(p["kv"]), "U": torch.empty(own, dtype=torch.float64).pin_memory()}
            h2d += sum(pinned[c][k].numel() * pinned[c][k].element_size() for k in ("coords", "n1", "n2", "kd", "kv"))
            d2h += own * 8 + 8

        def e2e_step():
            for c in cases:
                p, h = prob[c], pinned[c]
                s = p["solver"]
                s.mesh.coords[s.coord_lo:s.coord_hi].copy_(h["coords"], non_blocking=True)
                s.mesh.n1.copy_(h["n1"], non_blocking=True)
                s.mesh.n2.copy_(h["n2"], non_blocking=True)
                kd = h["kd"].to(dev, non_blocking=True)
                kv = h["kv"].to(dev, non_blocking=True)
                K = s.assemble(fs.E_mod, fs.A, fs.I)
                r = s.load_case(K, kd, kv, react_dofs=p["react"], rtol=RTOL, precond=args.precond, gather_U=False,
                                maxit=MAXIT)
                lo = K.row_offset
                h["U"].copy_(r["U"][lo:lo + K.n_rows], non_blocking=True)
                torch.cuda.synchronize()
        api = ("DistributedSolver.assemble/load_case on pinned host arrays: each rank copies in its share of the mesh "
               "(its elements, its nodes + halo) and the BCs, and copies out its rows of U (bytes are per rank)")
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = total_dof / (ms_e2e * 1e-3) / 1e6

    # ---- roofline of the dominant kernel, measured live (CUDA events around it on its stream) in the timed steps
    peak, peak_src = measured_peaks()
    k_ms, k_n, k_bytes = prof[0], int(prof[1]), prof[2]
    roof = None
    if k_n:
        ach = k_bytes / (k_ms * 1e-3) / 1e9
        amg = info[cases[-1]]["precond"] == "amg"
        kname = ("pcg_amg_kernel (one persistent cooperative launch per solve: CG recurrences + the whole multigrid "
                 "V-cycle -- TMA-pipelined sweeps over the symmetric 3x3 block view of every level, restriction, "
                 "prolongation, 3x3-block Jacobi smoothing -- + fused dots; algorithmic bytes per iteration = what "
                 "these phases must stream once per use, myc_amg_bytes_per_iteration, DESIGN.md section 4)") if amg else \
            "pcg_fused_kernel (one persistent launch per solve; bytes = (its+1)*(52/9 nnz + 20 n) + its*(96|120|124|148) n)"
        traffic, traffic_src = (None, "multi-GPU run") if world > 1 else \
            ncu_traffic(f"{'pcg_amg_kernel' if amg else 'pcg_fused_kernel'}@{args.grid}Y", info["Y"]["iterations"])
        roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "frac_of_nominal_8TBs": ach / NOMINAL_HBM_GBS, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "avg_launch_us": k_ms / k_n * 1e3, "launches_sampled": k_n,
                "algorithmic_bytes_per_launch": k_bytes / k_n, "share_of_step": k_ms / ms_total if ms_total else None,
                "note": "per-rank operator and kernel time of this rank (rank 0)"}

    line = {
        "metric": "assemble+solve throughput", "value": value, "unit": "MDOF/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_dof),
        "ms_assemble": split["ms_assemble"], "ms_setup": split["ms_setup"], "ms_solve": split["ms_solve"],
        "e2e": {"value": e2e_value, "unit": "MDOF/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "api": api},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "load_cases": info,
    }
    # free the benchmark specimens before the one-shot records
    for c in cases:
        prob[c].clear()
    prob.clear()
    torch.cuda.empty_cache()

    # ---- one-shot records ------------------------------------------------------------------------
    def one_shot(shape, case, parity):
        from mycelium_fea_project_b200.synth import synth_network
        coords, n1, n2 = synth_network(*shape)
        p = make_problem(coords, n1, n2, case)
        solve_problem(p, gather_U=parity, rtol=RTOL_ONE_SHOT)   # warm (allocations, IPC mappings)
        p["last"] = None                                        # hand K / U back to the caching allocator: the timed pass must
        barrier()                                               # not pay a second multi-GB cudaMalloc next to the first
        t0 = time.perf_counter()
        r = solve_problem(p, gather_U=parity, rtol=RTOL_ONE_SHOT)
        barrier()
        wall = time.perf_counter() - t0
        rec = {"grid": list(shape), "n_dof": p["n_dof"], "load_case": case, "solver": f"{r['precond']}-PCG", "rtol": RTOL_ONE_SHOT,
               "iterations": r["iterations"], "ms_assemble": max_over_ranks(r["ms_assemble"]),
               "ms_setup": max_over_ranks(r["ms_setup"]), "ms_solve": max_over_ranks(r["ms_solve"]),
               "ms_wall": max_over_ranks(wall * 1e3), "relres": r["relres"], "true_relres": true_relres(p),
               "total_force": r["total_force"]}
        rec["us_per_iteration"] = rec["ms_solve"] * 1e3 / max(r["iterations"], 1)
        rec["mdof_per_s"] = p["n_dof"] / (rec["ms_assemble"] + rec["ms_setup"] + rec["ms_solve"]) / 1e3
        if parity and world > 1:
            U_dist = p["last"][1]["U"]
            if rank == 0:                                       # the same solve on ONE GPU, private context
                c1 = dv.Context(local)
                try:
                    m1 = dv.DeviceMesh.from_host(coords, n1, n2)
                    K1 = dv.assemble(c1, m1, fs.E_mod, fs.A, fs.I)
                    s1 = dv.apply_dirichlet(c1, K1, p["kd_d"], p["kv_d"], precond=args.precond)
                    x1, it1, _ = dv.pcg(c1, K1, s1, precond=args.precond, rtol=RTOL_ONE_SHOT, maxit=MAXIT)
                    U1 = dv.merge_solution(c1, K1, s1, x1)
                    F1 = dv.spmv(c1, K1, U1)
                    f1 = dv.gather_sum(c1, F1, p["react_d"])
                    rec["parity"] = {"relL2_U_vs_1gpu": float((torch.linalg.norm(U_dist - U1) / torch.linalg.norm(U1)).item()),
                                     "iterations_1gpu": it1, "total_force_1gpu": f1,
                                     "force_rel_diff": abs(r["total_force"] - f1) / abs(f1)}
                    del m1, K1, s1, x1, U1, F1
                finally:
                    c1.close()
            barrier()
        p.clear()
        torch.cuda.empty_cache()
        return rec

    if not args.no_strong:
        try:
            line["strong_4096"] = one_shot((args.strong_grid, args.strong_grid), "X", parity=True)
            line["strong_4096"]["note"] = (f"BASELINE configs[3]: the same {args.strong_grid}^2 mesh at every --gpus N (strong "
                                           "scaling); 1 warm + 1 timed pass; times are max over ranks")
        except Exception as exc:
            line["strong_4096"] = {"error": repr(exc)}
        if world == 8:
            try:
                line["config4_8192"] = {c: one_shot((8192, 8192), c, parity=False) for c in ("Y", "shear")}
            except Exception as exc:
                line["config4_8192"] = {"error": repr(exc)}
    if rank == 0 and world == 1:
        if not args.no_block6:
            line["roofline_block6"] = block6_roofline(ctx, dv, fs, lib, args, peak, peak_src)
        if not args.no_hbm_roofline:
            line["roofline_hbm"] = hbm_roofline(ctx, dv, fs, peak, peak_src, args.grid)
        if not args.no_ramp:
            try:
                line["ramp"] = ramp_record(fs, cpu=not args.no_cpu_baseline)
            except Exception as exc:
                line["ramp"] = {"error": repr(exc)}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        md.shutdown(ctx)                 # unmap peers' buffers everywhere, barrier, then free: the order CUDA IPC asks for
        dist.destroy_process_group()


def block6_roofline(ctx, dv, fs, lib, args, peak, peak_src):
    """The block-Jacobi PCG (pcg_fused_kernel) on the 2048^2 Y load case: one launch, timed with CUDA events."""
    import torch
    coords, n1, n2 = specimen("Y", args.grid, 1)
    mesh = dv.DeviceMesh.from_host(coords, n1, n2)
    hi, lo = fs.grip_nodes(coords, GRIP, 1)
    kd, kv = fs.build_bc(hi, lo, DISP, -DISP, 1)
    lib.myc_profile_reset(ctx.h, 1)
    r = fs.analyze_load_case(mesh, kd, kv, react_dofs=3 * hi + 1, rtol=RTOL, precond="block6")
    prof = (C.c_double * 4)()
    lib.myc_profile_get(ctx.h, prof)
    lib.myc_profile_reset(ctx.h, 0)
    ach = prof[2] / (prof[0] * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": "pcg_fused_kernel<block6, sym3> (one persistent launch; bytes = (its+1)*(52/9 nnz + 20 n) + its*124 n)",
           "workload": f"synthetic {args.grid}x{args.grid} grid, Y load case", "iterations": r.iterations,
           "ms_solve": r.ms_solve, "us_per_iteration": prof[0] * 1e3 / max(r.iterations, 1), "achieved": ach, "peak": peak,
           "unit": "GB/s", "frac": ach / peak, "frac_of_nominal_8TBs": ach / NOMINAL_HBM_GBS, "peak_source": peak_src,
           "traffic": None, "total_force": r.total_force,
           "true_relres": dv.true_residual(ctx, r.K, r.system, r.x)}
    del mesh, r
    torch.cuda.empty_cache()
    return out


def hbm_roofline(ctx, dv, fs, peak, peak_src, N=2048):
    """CSR SpMV on the 2048^2 operator (1.27 GB, 10x L2), L2 flushed between launches."""
    import torch
    from mycelium_fea_project_b200.synth import synth_network
    coords, n1, n2 = synth_network(N)
    mesh = dv.DeviceMesh.from_host(coords, n1, n2)
    K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    asm = []
    for _ in range(5):
        ev[0].record(); K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I); ev[1].record(); ev[1].synchronize()
        asm.append(ev[0].elapsed_time(ev[1]))
    asm_ms = float(np.median(asm[1:]))
    asm_bytes = 9 * mesh.n_elem + 24 * mesh.n_nodes + 12 * K.nnz + 4 * (K.n_rows + 1)
    x = torch.randn(K.n_rows, dtype=torch.float64, device=ctx.device)
    y = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.device)
    for _ in range(3):
        dv.spmv(ctx, K, x, y)
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dv.spmv(ctx, K, x, y); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    nbytes = 12 * K.nnz + 20 * K.n_rows
    ach = nbytes / (ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(f"myc_spmv_tma_kernel@{N}", None)
    return {"bound": "hbm", "kernel": "myc_spmv_tma_kernel<TmCfgBlock3, EpiPlain> (y = K x on the CSR: per-warp TMA "
                                      "bulk-copy ring, node-block multiply/sum)", "workload": f"synthetic {N}x{N} grid",
            "n_rows": K.n_rows, "nnz": K.nnz, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "frac_of_nominal_8TBs": ach / NOMINAL_HBM_GBS,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src, "avg_launch_us": ms * 1e3,
            "algorithmic_bytes_per_launch": nbytes, "l2": "256 MiB flush write between launches",
            "assembly": {"ms": asm_ms, "algorithmic_bytes": asm_bytes, "achieved": asm_bytes / (asm_ms * 1e-3) / 1e9,
                         "unit": "GB/s", "frac": asm_bytes / (asm_ms * 1e-3) / 1e9 / peak,
                         "kernel": "assembly (edge_degree / scan / edge_place / segment_order / block_count / row_ptr / fill_staged kernels): mesh -> CSR, "
                                   "bytes = 9 n_elem + 24 n_nodes + 12 nnz + 4 (n_dof + 1) (the algorithmic minimum)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=2048)
    ap.add_argument("--strong-grid", type=int, default=4096)
    ap.add_argument("--precond", default=os.environ.get("MYC_PCG_PRECOND", "amg"),
                    choices=["amg", "jacobi", "block3", "block6", "block12"],
                    help="aggregation multigrid (default), or a (block-)Jacobi variant")
    ap.add_argument("--reference-budget", type=float, default=600.0,
                    help="--impl reference stops adding steps when the next one would pass this many seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-roofline", action="store_true")
    ap.add_argument("--no-block6", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-ramp", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
