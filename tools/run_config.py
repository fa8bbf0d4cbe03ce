"""Run one BASELINE.json configuration (square NxN synthetic network, strong scaling over the
launched ranks) and print one JSON line per load case: assembly time, solve time, iterations,
MDOF/s, final true residual.

  python tools/run_config.py --grid 2048 --cases Y
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config.py --grid 4096 --cases X,Y
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs, dist as md
from mycelium_fea_project_b200.synth import synth_network

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=2048)
ap.add_argument("--cases", default="Y")
ap.add_argument("--rtol", type=float, default=1e-10)
ap.add_argument("--maxit", type=int, default=600000)
ap.add_argument("--precond", default="jacobi")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = dv.Context.get(torch.device("cuda", local))
t0 = time.time()
coords, n1, n2 = synth_network(a.grid)
t_gen = time.time() - t0
n_dof = 3 * len(coords)
solver = md.DistributedSolver((coords, n1, n2), device=ctx.device) if world > 1 else None
mesh = solver.mesh if solver else dv.DeviceMesh.from_host(coords, n1, n2)
for case in a.cases.split(","):
    axis, comp = fs.LOAD_CASES[case]
    hi, lo = fs.grip_nodes(coords, 1.5, axis)
    kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, comp)
    react = 3 * hi + comp
    for rep in range(a.reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        if world > 1:
            K = solver.assemble(fs.E_mod, fs.A, fs.I)
            e[1].record()
            out = solver.load_case(K, kd, kv, react_dofs=react, rtol=a.rtol, precond=a.precond, maxit=a.maxit, gather_U=False)
            e[2].record(); e[2].synchronize()
            tr = solver.true_residual(K, out["system"], out["x"])
            its, rel, force, nnz = out["iterations"], out["relres"], out["total_force"], K.nnz
        else:
            fs.PCG_MAXIT = a.maxit
            K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
            e[1].record()
            r = fs.analyze_load_case(mesh, kd, kv, react_dofs=react, rtol=a.rtol, precond=a.precond, K=K)
            e[2].record(); e[2].synchronize()
            tr = dv.true_residual(ctx, r.K, r.system, r.x)
            its, rel, force, nnz = r.iterations, r.relres, r.total_force, K.nnz
        ms_a, ms_s = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        t = torch.tensor([ms_a, ms_s], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_a, ms_s = t.tolist()
        if rank == 0:
            print(json.dumps({"grid": a.grid, "n_gpus": world, "case": case, "rep": rep, "n_dof": n_dof, "nnz_rank0": nnz,
                              "precond": a.precond, "rtol": a.rtol, "t_assemble_ms": round(ms_a, 3),
                              "t_solve_ms": round(ms_s, 1), "iterations": its, "us_per_iter": round(1e3 * ms_s / max(its, 1), 2),
                              "MDOF_per_s": round(n_dof / ((ms_a + ms_s) * 1e-3) / 1e6, 4), "relres": rel, "true_relres": tr,
                              "total_force": force, "gen_s": round(t_gen, 1)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
