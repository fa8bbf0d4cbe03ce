"""Regenerate tests/golden/ from the reference tree (run in the build container only).

    python tests/golden/make_golden.py

1. Copies the reference's committed hand-made fixtures (DATA, not source):
   results/test_{I,X,t,y}/{nodes,elements}.csv + fea_results/*.csv, and the real
   snapshot results/sim_20251117_181147 (inputs as csv.gz, the two surviving golden
   outputs; active_elements.csv is packed to a bit array).
2. Imports the UNMODIFIED reference (oracle.ref_shim) and records what its own
   functions return on seeded inputs, so intermediate parity (K_e, CSR, BC sets, U)
   is pinned by reference outputs rather than by our restatement:
     ke_random.npz        bar_stiffness_bulk on 4096 random 3-D segments (+ edge cases)
     asm_synth64.npz      assemble_global_stiffness (verbatim 36-append loop) on the 64^2 grid
     asm_real.npz         same on sim_20251117_181147 (duplicate node pairs, z==0 zeros)
     solve_synth64.npz    BC sets of the step loop + solve_system U on the 64^2 grid
     solve_synth128.npz   same on the 128^2 grid, tol 1.5 (committed constants)
     ramp_real_step*.npz  U / reactions of ramp steps 1 and 5 of the real snapshot
"""
import gzip
import os
import shutil
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from mycelium_fea_project_b200.synth import synth_network  # noqa: E402

REF = ref_shim.REFERENCE_ROOT


def copy_fixtures():
    for name in ("test_I", "test_X", "test_t", "test_y"):
        src = os.path.join(REF, "results", name)
        dst = os.path.join(HERE, "ref_results", name)
        os.makedirs(os.path.join(dst, "fea_results"), exist_ok=True)
        for f in ("nodes.csv", "elements.csv"):
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        for f in ("stress_record.csv", "active_elements.csv", "node_displacements.csv",
                  "force_displacement.csv"):
            shutil.copyfile(os.path.join(src, "fea_results", f), os.path.join(dst, "fea_results", f))
    src = os.path.join(REF, "results", "sim_20251117_181147")
    dst = os.path.join(HERE, "ref_results", "sim_20251117_181147")
    os.makedirs(os.path.join(dst, "fea_results"), exist_ok=True)
    for f in ("nodes.csv", "elements.csv"):
        with open(os.path.join(src, f), "rb") as fi, gzip.open(os.path.join(dst, f + ".gz"), "wb") as fo:
            fo.write(fi.read())
    shutil.copyfile(os.path.join(src, "fea_results", "force_displacement.csv"),
                os.path.join(dst, "fea_results", "force_displacement.csv"))
    act = pd.read_csv(os.path.join(src, "fea_results", "active_elements.csv"))
    steps = act["step"].values
    bits = act.drop(columns="step").values.astype(bool)
    np.savez_compressed(os.path.join(dst, "fea_results", "active_elements.npz"),
                        packed=np.packbits(bits, axis=1), n_elems=bits.shape[1], step=steps)


def elems_df(n1, n2):
    return pd.DataFrame({"elem_id": np.arange(len(n1)), "n1": np.asarray(n1, dtype=np.int64),
                         "n2": np.asarray(n2, dtype=np.int64)})


def ref_bc(ref, top, bot, dy_top, dy_bot):
    """The dict construction of the reference's step loop (it is inline in fea_solver(),
    src/fea_solver_no_plotting.py:223-245, so it is re-enacted here with the same
    statements to obtain the reference's ordering)."""
    disp = {}
    for n in top:
        disp.update({3 * n + 0: 0.0, 3 * n + 1: dy_top, 3 * n + 2: 0.0})
    for n in bot:
        disp.update({3 * n + 0: 0.0, 3 * n + 1: dy_bot, 3 * n + 2: 0.0})
    kd = np.array(list(disp.keys()))
    kv = np.array([disp[k] for k in kd])
    return kd, kv


def gen_vectors():
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(20261018)

    # --- K_e on random 3-D segments, plus degenerate / tiny / axis-aligned cases
    p1 = rng.standard_normal((4096, 3)) * 0.5
    p2 = p1 + rng.standard_normal((4096, 3)) * np.array([0.05, 0.05, 0.02])
    p2[0] = p1[0]                                  # zero length -> L clamp
    p2[1] = p1[1] + np.array([1e-13, 0, 0])        # below clamp
    p2[2] = p1[2] + np.array([0.05, 0, 0])         # axis aligned
    p2[3] = p1[3] + np.array([0, -0.05, 0])
    p2[4] = p1[4] + np.array([0, 0, 0.05])
    p2[5] = p1[5] + np.array([3e-7, 4e-7, 0])
    K, L = ref.bar_stiffness_bulk(p1, p2)
    np.savez_compressed(os.path.join(HERE, "ke_random.npz"), p1=p1, p2=p2, K=K, L=L)

    # --- verbatim assembly on the 64^2 synthetic grid
    c, n1, n2 = synth_network(64)
    active = np.ones(len(n1), dtype=bool)
    Kc = ref.assemble_global_stiffness(c, elems_df(n1, n2), active)
    active2 = active.copy()
    active2[rng.random(len(n1)) < 0.3] = False      # 30 % failed elements
    Kc2 = ref.assemble_global_stiffness(c, elems_df(n1, n2), active2)
    np.savez_compressed(os.path.join(HERE, "asm_synth64.npz"), indptr=Kc.indptr, indices=Kc.indices,
                        data=Kc.data, active2=active2, indptr2=Kc2.indptr, indices2=Kc2.indices,
                        data2=Kc2.data)

    # --- BC + solve on the 64^2 grid (tol 0.5 so that most DOFs stay free)
    nodes = pd.DataFrame({"node_id": np.arange(len(c)), "x": c[:, 0], "y": c[:, 1], "z": c[:, 2]})
    for N, tol in ((64, 0.5), (128, 1.5)):
        c, n1, n2 = synth_network(N)
        nodes = pd.DataFrame({"node_id": np.arange(len(c)), "x": c[:, 0], "y": c[:, 1], "z": c[:, 2]})
        ymin, ymax = c[:, 1].min(), c[:, 1].max()
        top = nodes.loc[np.abs(nodes["y"] - ymax) < tol, "node_id"].values.astype(int)
        bot = nodes.loc[np.abs(nodes["y"] - ymin) < tol, "node_id"].values.astype(int)
        kd, kv = ref_bc(ref, top, bot, 0.02, -0.02)
        Kc = ref.assemble_global_stiffness(c, elems_df(n1, n2), np.ones(len(n1), dtype=bool))
        U = ref.solve_system(Kc, kd, kv)
        F = Kc @ U
        np.savez_compressed(os.path.join(HERE, f"solve_synth{N}.npz"), tol=tol, top=top, bot=bot,
                            known_dofs=kd, known_vals=kv, U=U,
                            total_force=F[[3 * n + 1 for n in top]].sum())

    # --- the real snapshot: assembly (duplicate pairs, explicit zeros) and two ramp steps
    src = os.path.join(REF, "results", "sim_20251117_181147")
    nodes = pd.read_csv(os.path.join(src, "nodes.csv"))
    elems = pd.read_csv(os.path.join(src, "elements.csv"))
    c = nodes[["x", "y", "z"]].values
    Kc = ref.assemble_global_stiffness(c, elems, np.ones(len(elems), dtype=bool))
    np.savez_compressed(os.path.join(HERE, "asm_real.npz"), indptr=Kc.indptr, indices=Kc.indices,
                        data=Kc.data)
    ymin, ymax = c[:, 1].min(), c[:, 1].max()
    top = nodes.loc[np.abs(nodes["y"] - ymax) < ref.GRIP_LENGTH, "node_id"].values.astype(int)
    bot = nodes.loc[np.abs(nodes["y"] - ymin) < ref.GRIP_LENGTH, "node_id"].values.astype(int)
    for step in (1, 5):
        f = step / (ref.N_STEPS - 1)
        kd, kv = ref_bc(ref, top, bot, ref.DISPLACEMENT_MAX * f, -ref.DISPLACEMENT_MAX * f)
        U = ref.solve_system(Kc, kd, kv)
        F = Kc @ U
        np.savez_compressed(os.path.join(HERE, f"ramp_real_step{step}.npz"), top=top, bot=bot,
                            known_dofs=kd, known_vals=kv, U=U,
                            total_force=F[[3 * n + 1 for n in top]].sum())


if __name__ == "__main__":
    if not ref_shim.available():
        sys.exit("reference tree not present; goldens can only be regenerated in the build container")
    copy_fixtures()
    gen_vectors()
    print("golden fixtures written to", HERE)
