"""Quick device timings (not the bench): assembly, SpMV GB/s, PCG per-iteration time."""
import argparse
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs
from mycelium_fea_project_b200.synth import synth_network


def ev_time(fn, reps, flush=None):
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="512,2048")
    ap.add_argument("--maxit", type=int, default=2_000_000)
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--precond", default="jacobi")
    ap.add_argument("--solve", type=int, default=1)
    a = ap.parse_args()
    ctx = dv.Context.get()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for N in [int(g) for g in a.grids.split(",")]:
        t0 = time.time()
        coords, n1, n2 = synth_network(N)
        t_gen = time.time() - t0
        mesh = dv.DeviceMesh.from_host(coords, n1, n2)
        K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
        torch.cuda.synchronize()
        asm_med, asm_min = ev_time(lambda: dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I), 5, flush)
        n_rows, nnz = K.n_rows, K.nnz
        x = torch.randn(n_rows, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        dv.spmv(ctx, K, x, y)
        sp_med, sp_min = ev_time(lambda: dv.spmv(ctx, K, x, y), 20, flush)
        sp_bytes = 12 * nnz + 20 * n_rows
        asm_bytes = 9 * mesh.n_elem + 24 * mesh.n_nodes + 12 * nnz + 4 * (n_rows + 1)
        out = {"N": N, "n_dof": n_rows, "nnz": nnz, "gen_s": round(t_gen, 2),
               "assemble_ms": round(asm_med, 3), "assemble_GBs_algmin": round(asm_bytes / asm_med / 1e6, 1),
               "spmv_ms": round(sp_med, 4), "spmv_ms_min": round(sp_min, 4),
               "spmv_GBs": round(sp_bytes / sp_med / 1e6, 1)}
        if a.solve:
            hi, lo = fs.grip_nodes(coords, 1.5, 1)
            kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, 1)
            sysd = dv.apply_dirichlet(ctx, K, torch.from_numpy(kd).cuda(), torch.from_numpy(kv).cuda(),
                                      precond=a.precond)
            dv.pcg(ctx, K, sysd, precond=a.precond, rtol=a.rtol, maxit=20, raise_on_maxit=False)   # warm-up
            torch.cuda.synchronize()
            t0 = time.time()
            xs, it, rel = dv.pcg(ctx, K, sysd, precond=a.precond, rtol=a.rtol, maxit=a.maxit, raise_on_maxit=False)
            torch.cuda.synchronize()
            dt = time.time() - t0
            tr = dv.true_residual(ctx, K, sysd, xs)
            it_bytes = 12 * nnz + 92 * n_rows          # CSR-sweep convention (the fused kernel streams less)
            out.update({"pcg_iters": it, "pcg_s": round(dt, 3), "us_per_iter": round(dt / max(it, 1) * 1e6, 2),
                        "iter_GBs": round(it_bytes * it / dt / 1e9, 1), "relres": rel, "true_relres": tr,
                        "precond": a.precond})
        print(json.dumps(out), flush=True)
        del K, mesh, x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
