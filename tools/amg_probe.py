"""Quick device probe of the AMG-PCG path: hierarchy, iterations, time, true residual (not the bench)."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs
from mycelium_fea_project_b200.synth import synth_network

ap = argparse.ArgumentParser()
ap.add_argument("--grids", default="512,2048")
ap.add_argument("--case", default="Y")
ap.add_argument("--rtol", type=float, default=1e-10)
ap.add_argument("--compare", default="", help="also solve with this preconditioner (e.g. block6)")
a = ap.parse_args()
ctx = dv.Context.get()
for N in [int(g) for g in a.grids.split(",")]:
    coords, n1, n2 = synth_network(N)
    axis, comp = fs.LOAD_CASES[a.case]
    hi, lo = fs.grip_nodes(coords, 1.5, axis)
    kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, comp)
    mesh = dv.DeviceMesh.from_host(coords, n1, n2)
    K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
    kdd, kvd = torch.from_numpy(kd).cuda(), torch.from_numpy(kv).cuda()
    torch.cuda.synchronize()
    t0 = time.time()
    sysd = dv.apply_dirichlet(ctx, K, kdd, kvd, precond="amg")
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    levels, setup_ms = dv.amg_levels(ctx) if sysd.amg_levels else ([], 0.0)
    out = {"N": N, "n_dof": K.n_rows, "precond": sysd.precond, "levels": levels, "amg_setup_ms": setup_ms,
           "dirichlet_plus_setup_ms": t_setup * 1e3}
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        x, it, rel = dv.pcg(ctx, K, sysd, precond="amg", rtol=a.rtol, raise_on_maxit=False, maxit=5000)
        torch.cuda.synchronize()
        dt = time.time() - t0
    out.update({"iters": it, "solve_ms": dt * 1e3, "us_per_iter": dt / max(it, 1) * 1e6, "relres": rel,
                "true_relres": dv.true_residual(ctx, K, sysd, x)})
    if a.compare:
        s2 = dv.apply_dirichlet(ctx, K, kdd, kvd, precond=a.compare)
        torch.cuda.synchronize(); t0 = time.time()
        x2, it2, rel2 = dv.pcg(ctx, K, s2, precond=a.compare, rtol=a.rtol)
        torch.cuda.synchronize(); dt2 = time.time() - t0
        out.update({"cmp": a.compare, "cmp_iters": it2, "cmp_solve_ms": dt2 * 1e3,
                    "relL2_vs_cmp": float((torch.linalg.norm(x - x2) / torch.linalg.norm(x2)).item())})
    print(json.dumps(out), flush=True)
    del K, mesh, sysd, x
    torch.cuda.empty_cache()
