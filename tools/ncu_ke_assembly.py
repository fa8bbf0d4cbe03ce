"""K_e generation and CSR assembly on one synthetic grid: event timings (plain run) or a single
un-timed round for an ncu capture (MYC_NCU=1).

    python tools/ncu_ke_assembly.py --grid 2048                 # JSON line with achieved GB/s
    MYC_NCU=1 ncu --set full --clock-control none --import-source on \
        -k regex:'ke_batch|edge_|rs_|block_count|row_ptr|fill_kernel' -c 24 -o gpurun_out/prof_ke_asm \
        python tools/ncu_ke_assembly.py --grid 2048

Algorithmic bytes: K_e 345 B/element (57 read + 288 written);
assembly 9 n_elem + 24 n_nodes + 12 nnz + 4 (n_dof + 1).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mycelium_fea_project_b200 import device as dv, fea_solver as fs
from mycelium_fea_project_b200.synth import synth_network


def ev(fn, reps, flush):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=2048)
    a = ap.parse_args()
    ctx = dv.Context.get()
    coords, n1, n2 = synth_network(a.grid)
    mesh = dv.DeviceMesh.from_host(coords, n1, n2)
    p1s = mesh.coords[mesh.n1.long()].contiguous()
    p2s = mesh.coords[mesh.n2.long()].contiguous()
    torch.cuda.synchronize()
    if os.environ.get("MYC_NCU") == "1":
        dv.bar_stiffness(ctx, p1s, p2s, fs.E_mod, fs.A, fs.I)
        dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
        torch.cuda.synchronize()
        return
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.device)
    n_elem, n_nodes = mesh.n_elem, mesh.n_nodes
    Ke = torch.empty((n_elem, 6, 6), dtype=torch.float64, device=ctx.device)
    L = torch.empty((n_elem,), dtype=torch.float64, device=ctx.device)
    from mycelium_fea_project_b200._lib import lib, check

    def ke():
        check(ctx.h, lib.myc_bar_stiffness_bulk(ctx.h, p1s.data_ptr(), p2s.data_ptr(), n_elem, float(fs.E_mod),
                                                fs.A, fs.I, Ke.data_ptr(), L.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    for _ in range(3):
        ke()
    ke_ms, ke_min = ev(ke, 10, flush)
    K = dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
    nnz, n_dof = K.nnz, K.n_rows
    del K
    for _ in range(2):
        dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I)
    asm_ms, asm_min = ev(lambda: dv.assemble(ctx, mesh, fs.E_mod, fs.A, fs.I), 5, flush)
    ke_bytes = 345 * n_elem                       # 48 coords + 8 ids + 1 active read, 288 written
    ke_moved = (48 + 288 + 8) * n_elem            # what THIS entry point moves: p1s, p2s in; K_e, L out
    asm_bytes = 9 * n_elem + 24 * n_nodes + 12 * nnz + 4 * (n_dof + 1)
    print(json.dumps({
        "grid": a.grid, "n_elem": n_elem, "n_nodes": n_nodes, "n_dof": n_dof, "nnz": nnz,
        "ke_ms": ke_ms, "ke_ms_min": ke_min, "ke_algorithmic_GBs": ke_bytes / ke_ms / 1e6,
        "ke_moved_GBs": ke_moved / ke_ms / 1e6, "ke_flops": 70 * n_elem, "ke_GFLOPs": 70 * n_elem / ke_ms / 1e6,
        "assemble_ms": asm_ms, "assemble_ms_min": asm_min, "assemble_algorithmic_GBs": asm_bytes / asm_ms / 1e6,
        "l2": "256 MiB flush write between launches"}))


if __name__ == "__main__":
    main()
