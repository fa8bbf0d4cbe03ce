"""ncu target: assemble the NxN grid once, then a few plain CSR SpMVs and PCG iterations."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs
from mycelium_fea_project_b200.synth import synth_network

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = dv.Context.get()
coords, n1, n2 = synth_network(N)
mesh = dv.DeviceMesh.from_host(coords, n1, n2)
K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
x = torch.randn(K.n_rows, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for _ in range(iters):
    dv.spmv(ctx, K, x, y)
hi, lo = fs.grip_nodes(coords, 1.5, 1)
kd, kv = fs.build_bc(hi, lo, 0.02, -0.02, 1)
sysd = dv.apply_dirichlet(ctx, K, torch.from_numpy(kd).cuda(), torch.from_numpy(kv).cuda())
dv.pcg(ctx, K, sysd, rtol=1e-10, maxit=iters, raise_on_maxit=False)
torch.cuda.synchronize()
print("ok", K.n_rows, K.nnz)
