"""ncu target: ONE complete load case of the bench workload (512^2 grid, Y case, block-Jacobi PCG to
rtol 1e-10) = one launch of the persistent pcg_fused_kernel, for the steady-state DRAM traffic of the
dominant kernel (the 3-iteration capture in profiles/r1_bench_launches.md is cold-cache).

    ncu --set full --clock-control none --import-source on -k regex:pcg_fused -c 1 \
        -o gpurun_out/prof_fused_solve python tools/ncu_fused_solve.py [grid]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs
from mycelium_fea_project_b200.synth import synth_network

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
coords, n1, n2 = synth_network(N)
mesh = dv.DeviceMesh.from_host(coords, n1, n2)
hi, lo = fs.grip_nodes(coords, 1.5, 1)
kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, 1)
res = fs.analyze_load_case(mesh, kd, kv, react_dofs=3 * hi + 1, rtol=1e-10)
torch.cuda.synchronize()
n, nnz, its = res.K.n_rows, res.K.nnz, res.iterations
vec = {"jacobi": 96, "block3": 120, "block6": 124, "block12": 148}[fs.PCG_PRECOND]   # B per row per iteration
alg = (its + 1) * (52 / 9 * nnz + 20 * n) + its * vec * n
print(f"ok precond={fs.PCG_PRECOND} grid={N} n_dof={n} nnz={nnz} iterations={its} relres={res.relres:.3e} force={res.total_force:.6e} "
      f"algorithmic_bytes_per_launch={alg:.6e} ms_solve={res.ms_solve:.2f}")
