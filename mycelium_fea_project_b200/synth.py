"""Synthetic grid-occupancy mycelium networks (the benchmark inputs of BASELINE.json).

Not part of the reference: SURVEY.md section 8(d) defines this generator so that every
size named in BASELINE.json ("synthetic NxN mycelium occupancy grid") is reproducible
from a seed.  Output is the reference's snapshot schema (nodes.csv: node_id,x,y,z;
elements.csv: elem_id,n1,n2 -- writers src/mycelium_sim_2D.py:723-727).
"""
from __future__ import annotations

import os

import numpy as np

SEGMENT_LENGTH = 0.05      # mm, the reference's hyphal segment length (src/mycelium_sim_2D.py:23)
OCCUPANCY = 2.0 / 3.0


def synth_network(n_rows, n_cols=None, seed=0, p=OCCUPANCY, h=SEGMENT_LENGTH):
    """Site-percolation network on an n_rows x n_cols lattice.

    Returns (coords (n_nodes,3) f64, n1 (n_elem,) i32, n2 (n_elem,) i32).
    Nodes are the occupied sites numbered row-major (y outer, x inner), jittered by
    U(-0.2h, 0.2h) in x and y, z = 0.  Elements are the bonds to occupied right
    neighbours (row-major) followed by the bonds to occupied up neighbours; n1 < n2.
    Isolated sites are kept (they give empty CSR rows).
    """
    n_cols = n_rows if n_cols is None else n_cols
    rng = np.random.default_rng(seed)
    occ = rng.random((n_rows, n_cols)) < p
    n_nodes = int(occ.sum())
    ids = np.full((n_rows, n_cols), -1, dtype=np.int64)
    ids[occ] = np.arange(n_nodes)
    iy, ix = np.nonzero(occ)
    jitter = rng.uniform(-0.2 * h, 0.2 * h, size=(n_nodes, 2))
    coords = np.zeros((n_nodes, 3))
    coords[:, 0] = ix * h + jitter[:, 0]
    coords[:, 1] = iy * h + jitter[:, 1]
    right = occ[:, :-1] & occ[:, 1:]
    up = occ[:-1, :] & occ[1:, :]
    n1 = np.concatenate([ids[:, :-1][right], ids[:-1, :][up]])
    n2 = np.concatenate([ids[:, 1:][right], ids[1:, :][up]])
    return coords, n1.astype(np.int32), n2.astype(np.int32)


def write_snapshot(results_dir, coords, n1, n2, binary_sidecar=True):
    """Write nodes.csv / elements.csv in the reference layout, plus (optionally) a
    ``mesh.npz`` side-car that the drop-in loader prefers for >= 2048^2 meshes where CSV
    parsing would dominate."""
    import pandas as pd
    os.makedirs(results_dir, exist_ok=True)
    n_nodes = coords.shape[0]
    pd.DataFrame({"node_id": np.arange(n_nodes), "x": coords[:, 0], "y": coords[:, 1],
                  "z": coords[:, 2]}).to_csv(os.path.join(results_dir, "nodes.csv"), index=False)
    pd.DataFrame({"elem_id": np.arange(len(n1)), "n1": n1, "n2": n2}).to_csv(
        os.path.join(results_dir, "elements.csv"), index=False)
    if binary_sidecar:
        np.savez(os.path.join(results_dir, "mesh.npz"), coords=coords, n1=n1, n2=n2)
