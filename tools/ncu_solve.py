"""ncu target: ONE complete load case (Y) on a synthetic grid = one launch of the persistent solver kernel
(pcg_amg_kernel for --precond amg, pcg_fused_kernel for the block-Jacobi family).

    python tools/ncu_solve.py 2048 amg > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:pcg_amg -c 1 \
        -o gpurun_out/prof_amg_2048 python tools/ncu_solve.py 2048 amg

Prints the iteration count, the solve time, and the algorithmic bytes of the launch as the library counts
them for bench.py's roofline (myc_profile_get), so that dram__bytes of the capture can be set against them.
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs
from mycelium_fea_project_b200._lib import lib
from mycelium_fea_project_b200.synth import synth_network

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
precond = sys.argv[2] if len(sys.argv) > 2 else "amg"
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else fs.PCG_MAXIT
ctx = dv.Context.get()
coords, n1, n2 = synth_network(N)
mesh = dv.DeviceMesh.from_host(coords, n1, n2)
hi, lo = fs.grip_nodes(coords, 1.5, 1)
kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, 1)
fs.PCG_MAXIT = maxit
lib.myc_profile_reset(ctx.h, 1)
try:
    res = fs.analyze_load_case(mesh, kd, kv, react_dofs=3 * hi + 1, rtol=1e-10, precond=precond)
    its, relres = res.iterations, res.relres
except fs.MyceliumFeaError as exc:          # a capped capture (maxit) is fine for profiling
    its, relres, res = maxit, float("nan"), None
    print("note:", exc)
torch.cuda.synchronize()
prof = (C.c_double * 4)()
lib.myc_profile_get(ctx.h, prof)
print(json.dumps({"grid": N, "precond": precond, "iterations": its, "relres": relres,
                  "kernel_ms": prof[0], "launches": int(prof[1]), "algorithmic_bytes_per_launch": prof[2] / max(prof[1], 1),
                  "algorithmic_bytes_per_iteration": prof[2] / max(prof[1], 1) / (its + 1),
                  "GBs": prof[2] / max(prof[0], 1e-9) / 1e6,
                  "ms_assemble": getattr(res, "ms_assemble", None), "ms_setup": getattr(res, "ms_setup", None),
                  "ms_solve": getattr(res, "ms_solve", None)}))
