"""Assemble the NxN synthetic grid on the GPU and dump the CSR to a flat binary (for tools/spmv_bench.cu)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs
from mycelium_fea_project_b200.synth import synth_network

N, path = int(sys.argv[1]), sys.argv[2]
ctx = dv.Context.get()
coords, n1, n2 = synth_network(N)
K = dv.assemble(ctx, dv.DeviceMesh.from_host(coords, n1, n2), fs.E_mod, fs.A, fs.I)
with open(path, "wb") as f:
    np.array([K.n_rows, K.nnz], dtype=np.int64).tofile(f)
    K.row_ptr.cpu().numpy().tofile(f)
    K.col_idx.cpu().numpy().tofile(f)
    K.val.cpu().numpy().tofile(f)
print("dumped", N, K.n_rows, K.nnz, path)
