"""Multi-GPU diagnostic: X and Y specimens of the bench, true residuals right after each solve and
again after the other mesh's solve (different halo plan installed)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs, dist as md
from mycelium_fea_project_b200.synth import synth_network

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = dv.Context.get(torch.device("cuda", local))
S = {}
for c in ("X", "Y"):
    coords, n1, n2 = synth_network(grid, grid * world) if c == "Y" else synth_network(grid * world, grid)
    axis, comp = fs.LOAD_CASES[c]
    hi, lo = fs.grip_nodes(coords, 1.5, axis)
    kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, comp)
    S[c] = (md.DistributedSolver((coords, n1, n2), device=ctx.device), kd, kv)
res = {}
for rep in range(2):
    for c in ("X", "Y"):
        s, kd, kv = S[c]
        K = s.assemble(fs.E_mod, fs.A, fs.I)
        out = s.load_case(K, kd, kv, rtol=1e-10, gather_U=False, maxit=100000)
        tr = dv.true_residual(ctx, K, out["system"], out["x"])
        res[c] = (K, out)
        if rank == 0:
            print(f"rep {rep} {c}: it={out['iterations']} relres={out['relres']:.3e} true(right after)={tr:.3e}", flush=True)
    for c in ("X", "Y"):
        K, out = res[c]
        tr_other_plan = dv.true_residual(ctx, K, out["system"], out["x"])
        S[c][0]._install_plan()
        tr_own_plan = dv.true_residual(ctx, K, out["system"], out["x"])
        if rank == 0:
            print(f"rep {rep} {c}: true(last installed plan)={tr_other_plan:.3e} true(own plan)={tr_own_plan:.3e}", flush=True)
md.shutdown(ctx)
dist.destroy_process_group()
