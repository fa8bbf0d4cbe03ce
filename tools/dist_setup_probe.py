"""Multi-GPU probe: alternate the X and Y weak-scaling specimens (two DistributedSolver objects on one context, like
bench.py) and print where the time before the PCG goes (torch.distributed.run, one rank per GPU)."""
import json
import os
import sys
import time

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
ctx = dv.Context.get(torch.device("cuda", local))
S = {}
for c in ("X", "Y"):
    coords, n1, n2 = synth_network(grid * world, grid) if c == "X" else synth_network(grid, grid * world)
    axis, comp = fs.LOAD_CASES[c]
    hi, lo = fs.grip_nodes(coords, 1.5, axis)
    kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, comp)
    S[c] = (md.DistributedSolver((coords, n1, n2), device=ctx.device), torch.from_numpy(kd).to(ctx.device),
            torch.from_numpy(kv).to(ctx.device))
for rep in range(4):
    for c in ("X", "Y"):
        s, kd, kv = S[c]
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        K = s.assemble(fs.E_mod, fs.A, fs.I)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        sysd = dv.apply_dirichlet(ctx, K, kd, kv, precond="jacobi")
        torch.cuda.synchronize(); t2 = time.perf_counter()
        sysd = dv.apply_dirichlet(ctx, K, kd, kv, precond="amg")
        torch.cuda.synchronize(); t3 = time.perf_counter()
        _, setup_ms = dv.amg_levels(ctx)
        x, it, rel = dv.pcg(ctx, K, sysd, precond="amg", rtol=1e-10)
        torch.cuda.synchronize(); t4 = time.perf_counter()
        if rank == 0:
            print(json.dumps({"rep": rep, "case": c, "assemble_ms": (t1 - t0) * 1e3, "dirichlet_only_ms": (t2 - t1) * 1e3,
                              "dirichlet_plus_amg_ms": (t3 - t2) * 1e3, "amg_setup_events_ms": setup_ms,
                              "pcg_ms": (t4 - t3) * 1e3, "iterations": it}), flush=True)
md.shutdown(ctx)
dist.destroy_process_group()
