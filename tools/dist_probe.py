"""Multi-GPU probe of the multigrid PCG (launch with torch.distributed.run, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node N tools/dist_probe.py [grid] [case] [strong]
Weak specimen (grid x grid*N, like bench.py) unless `strong` is given (grid x grid on all ranks).  With a
-DMYC_AMG_TIMING build (MYC_LIB_PATH) every rank prints its per-phase times to stderr."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs, dist as md
from mycelium_fea_project_b200.synth import synth_network

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
case = sys.argv[2] if len(sys.argv) > 2 else "Y"
strong = len(sys.argv) > 3 and sys.argv[3] == "strong"
ctx = dv.Context.get(torch.device("cuda", local))
mult = 1 if strong else world
coords, n1, n2 = synth_network(grid * mult, grid) if case == "X" else synth_network(grid, grid * mult)
axis, comp = fs.LOAD_CASES[case]
hi, lo = fs.grip_nodes(coords, 1.5, axis)
kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, comp)
s = md.DistributedSolver((coords, n1, n2), device=ctx.device)
for rep in range(2):
    K = s.assemble(fs.E_mod, fs.A, fs.I)
    if rank != 0 or rep == 0:
        sys.stderr.flush()
    out = s.load_case(K, kd, kv, react_dofs=3 * hi + comp, rtol=1e-10, gather_U=False, maxit=20000)
    levels, _ = dv.amg_levels(ctx, detail=True) if out["precond"] == "amg" else ([], 0)
    if rank == 0:
        print(json.dumps({"rep": rep, "world": world, "grid": grid, "case": case, "strong": strong, "n_dof": 3 * len(coords),
                          "iterations": out["iterations"], "ms_setup": out["ms_setup"], "ms_solve": out["ms_solve"],
                          "us_per_iteration": out["ms_solve"] * 1e3 / max(out["iterations"], 1), "total_force": out["total_force"],
                          "levels": [(l["n_global"], l["replicated"]) for l in levels]}), flush=True)
md.shutdown(ctx)
dist.destroy_process_group()
