"""Snapshot I/O for large meshes.

The reference's formats are CSV: inputs ``nodes.csv`` / ``elements.csv`` (writers
src/mycelium_sim_2D.py:723-727, readers src/fea_solver.py:193-196) and four wide output CSVs with one
column per element / DOF (src/fea_solver.py:298-316).  Beyond ~2 M DOF the wide CSVs are unusable, so
the drop-in additionally understands binary side-cars: ``mesh.npz`` (coords, n1, n2) next to the input
CSVs and ``fea_results/records.npz`` (stress, active, disp, force_disp) instead of the output CSVs.
This module converts between the two losslessly.

    python -m mycelium_fea_project_b200.snapshot_io to-npz  results/sim_X     # nodes/elements.csv -> mesh.npz
    python -m mycelium_fea_project_b200.snapshot_io to-csv  results/sim_X     # mesh.npz -> nodes/elements.csv
    python -m mycelium_fea_project_b200.snapshot_io records-to-csv results/sim_X
"""
from __future__ import annotations

import os
import sys

import numpy as np


def csv_to_npz(results_dir):
    import pandas as pd
    nodes = pd.read_csv(os.path.join(results_dir, "nodes.csv"))          # parsed exactly like the reference does
    elems = pd.read_csv(os.path.join(results_dir, "elements.csv"))
    if not np.array_equal(nodes["node_id"].values, np.arange(len(nodes))):
        raise ValueError("node_id must equal the row index (the reference indexes coords by position, fea_solver.py:82)")
    out = os.path.join(results_dir, "mesh.npz")
    np.savez(out, coords=nodes[["x", "y", "z"]].values.astype(np.float64),
             n1=elems["n1"].values.astype(np.int32), n2=elems["n2"].values.astype(np.int32))
    return out


def npz_to_csv(results_dir):
    import pandas as pd
    z = np.load(os.path.join(results_dir, "mesh.npz"))
    c = z["coords"]
    pd.DataFrame({"node_id": np.arange(len(c)), "x": c[:, 0], "y": c[:, 1], "z": c[:, 2]}).to_csv(
        os.path.join(results_dir, "nodes.csv"), index=False)
    pd.DataFrame({"elem_id": np.arange(len(z["n1"])), "n1": z["n1"], "n2": z["n2"]}).to_csv(
        os.path.join(results_dir, "elements.csv"), index=False)


def records_to_csv(results_dir):
    """fea_results/records.npz -> the reference's four CSVs (same columns, same formatting)."""
    from .fea_solver import write_results
    fea_dir = os.path.join(results_dir, "fea_results")
    z = np.load(os.path.join(fea_dir, "records.npz"))
    rec = {"stress": list(z["stress"]), "active": list(z["active"]), "disp": list(z["disp"]),
           "force_disp": [list(r) for r in z["force_disp"]]}
    write_results(fea_dir, rec, z["stress"].shape[1])


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("to-npz", "to-csv", "records-to-csv"):
        sys.exit(__doc__)
    {"to-npz": csv_to_npz, "to-csv": npz_to_csv, "records-to-csv": records_to_csv}[sys.argv[1]](sys.argv[2])
