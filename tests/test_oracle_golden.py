"""Pin the oracle: reference goldens (committed outputs) and reference-generated vectors.

CPU only.  The reference tree is NOT needed (fixtures are committed); the tests that import
the live reference are skipped when /root/reference is absent.
"""
import filecmp
import gzip
import io
import os
import shutil

import numpy as np
import pandas as pd
import pytest

from oracle import fea_oracle as fo
from oracle import ref_shim
from mycelium_fea_project_b200.synth import synth_network

# constants the committed goldens were produced with
GOLDEN_CONSTS = {
    "test_X": dict(tol=0.5, disp_max=0.06, n_steps=40),
    "test_I": dict(tol=0.5, disp_max=0.06, n_steps=40),
    "test_y": dict(tol=0.5, disp_max=0.06, n_steps=100),
    "test_t": dict(tol=0.5, disp_max=2.0, n_steps=40),
}
CSVS = ("stress_record.csv", "active_elements.csv", "node_displacements.csv", "force_displacement.csv")


@pytest.mark.parametrize("name", sorted(GOLDEN_CONSTS))
def test_letter_fixture_bit_equal(name, golden_dir, tmp_path):
    src = os.path.join(golden_dir, "ref_results", name)
    for f in ("nodes.csv", "elements.csv"):
        shutil.copyfile(os.path.join(src, f), tmp_path / f)
    fo.fea_solver(str(tmp_path), **GOLDEN_CONSTS[name])
    for f in CSVS:
        assert filecmp.cmp(tmp_path / "fea_results" / f, os.path.join(src, "fea_results", f),
                           shallow=False), f"{name}/{f} differs from the committed golden"


def test_committed_constants_are_degenerate_on_test_X(golden_dir):
    """with GRIP_LENGTH=1.5 every node of test_X is in both grips."""
    nodes = pd.read_csv(os.path.join(golden_dir, "ref_results", "test_X", "nodes.csv"))
    c = nodes[["x", "y", "z"]].values
    top, bot = fo.grip_nodes(c, fo.GRIP_LENGTH)
    kd, kv = fo.build_bc(top, bot, 0.02, -0.02)
    assert len(kd) == 3 * len(c)           # zero free DOFs
    # node in both sets: first-insertion position, last value
    assert kd[1] == 22 and kv[1] == -0.02   # node 7 (hub): top position, bottom value


def _load_real(golden_dir):
    d = os.path.join(golden_dir, "ref_results", "sim_20251117_181147")
    nodes = pd.read_csv(io.BytesIO(gzip.open(os.path.join(d, "nodes.csv.gz")).read()))
    elems = pd.read_csv(io.BytesIO(gzip.open(os.path.join(d, "elements.csv.gz")).read()))
    return d, nodes, elems


@pytest.mark.slow
def test_real_snapshot_cascade(golden_dir):
    """results/sim_20251117_181147 with the committed constants: active cascade bit-equal
    over 40 steps, force-displacement to 1e-12 relative."""
    d, nodes, elems = _load_real(golden_dir)
    c = nodes[["x", "y", "z"]].values
    res = fo.fea_ramp(c, elems["n1"].values, elems["n2"].values)
    g = np.load(os.path.join(d, "fea_results", "active_elements.npz"))
    gold = np.unpackbits(g["packed"], axis=1)[:, :int(g["n_elems"])].astype(bool)
    assert np.array_equal(np.array(res.active), gold)
    fd = pd.read_csv(os.path.join(d, "fea_results", "force_displacement.csv"),
                     float_precision="round_trip").values
    mine = np.array(res.force_disp)
    assert mine.shape == fd.shape
    assert np.array_equal(mine[:, 0], fd[:, 0])
    assert np.abs(mine[:, 1] - fd[:, 1]).max() <= 1e-12 * np.abs(fd[:, 1]).max()


def test_ke_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "ke_random.npz"))
    K, L = fo.bar_stiffness_bulk(g["p1"], g["p2"])
    assert np.array_equal(L, g["L"])
    assert np.array_equal(K, g["K"])


def test_assembly_vectors(golden_dir):
    c, n1, n2 = synth_network(64)
    g = np.load(os.path.join(golden_dir, "asm_synth64.npz"))
    for act, sfx in ((np.ones(len(n1), bool), ""), (g["active2"], "2")):
        K = fo.assemble_global_stiffness(c, n1, n2, act)
        assert K.indptr.dtype == np.int32 and K.indices.dtype == np.int32
        assert np.array_equal(K.indptr, g["indptr" + sfx])
        assert np.array_equal(K.indices, g["indices" + sfx])
        assert np.array_equal(K.data, g["data" + sfx])
    # the literal append loop and the index-arithmetic restatement emit the same stream
    Kl = fo.assemble_global_stiffness_loop(c, n1, n2, np.ones(len(n1), bool))
    assert np.array_equal(Kl.indptr, g["indptr"]) and np.array_equal(Kl.data, g["data"])


def test_assembly_real_vectors(golden_dir):
    d, nodes, elems = _load_real(golden_dir)
    g = np.load(os.path.join(golden_dir, "asm_real.npz"))
    K = fo.assemble_global_stiffness(nodes[["x", "y", "z"]].values, elems["n1"].values,
                                     elems["n2"].values, np.ones(len(elems), bool))
    assert np.array_equal(K.indptr, g["indptr"])
    assert np.array_equal(K.indices, g["indices"])
    assert np.array_equal(K.data, g["data"])
    assert (K.data == 0).sum() > 80000       # explicit zeros are kept (z == 0 couplings)


@pytest.mark.parametrize("N", [64, 128])
def test_bc_and_solve_vectors(N, golden_dir):
    g = np.load(os.path.join(golden_dir, f"solve_synth{N}.npz"))
    c, n1, n2 = synth_network(N)
    top, bot = fo.grip_nodes(c, float(g["tol"]))
    assert np.array_equal(top, g["top"]) and np.array_equal(bot, g["bot"])
    kd, kv = fo.build_bc(top, bot, 0.02, -0.02)
    assert np.array_equal(kd, g["known_dofs"]) and np.array_equal(kv, g["known_vals"])
    K = fo.assemble_global_stiffness(c, n1, n2, np.ones(len(n1), bool))
    U = fo.solve_system(K, kd, kv)
    assert np.array_equal(U, g["U"])
    assert fo.reactions(K, U, top) == float(g["total_force"])


def test_pcg_baseline_matches_direct(golden_dir):
    g = np.load(os.path.join(golden_dir, "solve_synth64.npz"))
    c, n1, n2 = synth_network(64)
    K = fo.assemble_global_stiffness(c, n1, n2, np.ones(len(n1), bool))
    U, it, rel = fo.solve_system_pcg(K, g["known_dofs"], g["known_vals"], rtol=1e-12)
    assert rel <= 1e-12
    assert np.linalg.norm(U - g["U"]) <= 1e-8 * np.linalg.norm(g["U"])


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_live_reference_matches_oracle(tmp_path):
    """Belt and braces in the build container: run the imported reference itself."""
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(7)
    p1 = rng.standard_normal((500, 3)); p2 = p1 + 0.05 * rng.standard_normal((500, 3))
    Kr, Lr = ref.bar_stiffness_bulk(p1, p2)
    Ko, Lo = fo.bar_stiffness_bulk(p1, p2)
    assert np.array_equal(Kr, Ko) and np.array_equal(Lr, Lo)
    for k in ("E_mod", "A", "I", "N_STEPS", "DISPLACEMENT_MAX", "MAX_STRAIN", "GRIP_LENGTH"):
        assert getattr(ref, k) == getattr(fo, k)
    src = os.path.join(ref_shim.REFERENCE_ROOT, "results", "test_X")
    out = ref_shim.run_reference(src, tol=0.5, DISPLACEMENT_MAX=0.06, N_STEPS=40)
    try:
        for f in CSVS:
            assert filecmp.cmp(os.path.join(out, "fea_results", f),
                               os.path.join(src, "fea_results", f), shallow=False)
    finally:
        shutil.rmtree(out)


def test_petsc_style_c_port_agrees_with_direct_solve(golden_dir):
    """oracle/pcg_port.c (the multi-threaded restatement of the reference's PETSc path used as a CPU
    baseline) solves the same system: MatZeroRowsColumns-style elimination + Jacobi-PCG."""
    from oracle import pcg_port
    if not pcg_port.available():
        pytest.skip("oracle/_build/libpcg_port.so not built (run __graft_entry__.build())")
    g = np.load(os.path.join(golden_dir, "solve_synth64.npz"))
    c, n1, n2 = synth_network(64)
    K = fo.assemble_global_stiffness(c, n1, n2, np.ones(len(n1), bool))
    U, it, rel = pcg_port.solve_system_petsc_style(K, g["known_dofs"], g["known_vals"], rtol=1e-14, max_iters=100000)
    assert it < 100000 and rel <= 1e-14
    assert np.array_equal(U[g["known_dofs"]], g["known_vals"]) or np.allclose(U[g["known_dofs"]], g["known_vals"], rtol=1e-11)
    # PETSc's residual is relative to a b that contains the prescribed values, so the free part is
    # resolved less tightly than rtol suggests
    assert np.linalg.norm(U - g["U"]) <= 1e-6 * np.linalg.norm(g["U"])


@pytest.mark.parametrize("rows_per_block", [3, 6, 12])
def test_block_jacobi_checker_agrees_with_direct_solve(rows_per_block, golden_dir):
    """oracle.block_jacobi_pcg (the checker of the CUDA path's block preconditioners) solves the
    reduced system of src/fea_solver.py:118-125 to the same U_f as spsolve, in fewer iterations the
    larger the aligned blocks are; its block inverses are exact and symmetric."""
    g = np.load(os.path.join(golden_dir, "solve_synth64.npz"))
    from mycelium_fea_project_b200.synth import synth_network
    coords, n1, n2 = synth_network(64)
    K = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
    free, K_ff, F_f = fo.reduce_system(K, g["known_dofs"], g["known_vals"])
    x, it, rel = fo.block_jacobi_pcg(K_ff, F_f, free, rows_per_block, rtol=1e-12)
    ref = g["U"][free]
    assert rel <= 1e-12
    assert np.linalg.norm(x - ref) <= 1e-8 * np.linalg.norm(ref)
    _, it_point, _ = fo.jacobi_pcg(K_ff.tocsr(), F_f, rtol=1e-12)
    assert it < it_point
    labels, blocks = fo.aligned_block_inverses(K_ff, free, rows_per_block)
    assert len(blocks) == len(np.unique(labels))
    idx, inv = next(iter(blocks.values()))
    assert np.allclose(inv, inv.T, rtol=1e-10, atol=0.0) or inv.shape == (1, 1)
    assert np.allclose(inv @ K_ff.tocsr()[idx][:, idx].toarray(), np.eye(len(idx)), atol=1e-8)
