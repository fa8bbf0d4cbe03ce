"""CPU tests of host-side I/O helpers (no GPU compute)."""
import numpy as np

from mycelium_fea_project_b200.synth import synth_network, write_snapshot


def test_snapshot_roundtrip_csv_and_sidecar(tmp_path):
    from mycelium_fea_project_b200 import fea_solver as fs
    coords, n1, n2 = synth_network(32, 20, seed=9)
    write_snapshot(str(tmp_path / "s"), coords, n1, n2, binary_sidecar=True)
    c, a, b = fs.load_snapshot(str(tmp_path / "s"))            # prefers mesh.npz
    assert np.array_equal(c, coords) and np.array_equal(a, n1) and np.array_equal(b, n2)
    write_snapshot(str(tmp_path / "t"), coords, n1, n2, binary_sidecar=False)
    c, a, b = fs.load_snapshot(str(tmp_path / "t"))            # reference CSV schema, parsed like the reference
    # (pandas' default float parser, which the reference uses too, is not round-trip exact: <= 1 ulp)
    assert np.abs(c - coords).max() <= 1e-15 and np.array_equal(a, n1) and np.array_equal(b, n2)
    import pandas as pd
    nodes = pd.read_csv(tmp_path / "t" / "nodes.csv")
    elems = pd.read_csv(tmp_path / "t" / "elements.csv")
    assert list(nodes.columns) == ["node_id", "x", "y", "z"] and list(elems.columns) == ["elem_id", "n1", "n2"]
    assert np.array_equal(nodes["node_id"].values, np.arange(len(coords)))


def test_generator_is_seeded_and_matches_survey_counts():
    c1, a1, b1 = synth_network(64)
    c2, a2, b2 = synth_network(64)
    assert np.array_equal(c1, c2) and np.array_equal(a1, a2) and np.array_equal(b1, b2)
    assert (len(c1), len(a1)) == (2743, 3609)                 # SURVEY.md section 6 / BASELINE.md
    assert np.all(a1 < b1) and c1[:, 2].max() == 0.0
    c3, _, _ = synth_network(64, seed=1)
    assert len(c3) != len(c1) or not np.array_equal(c3, c1)


def test_converter_roundtrips_reference_snapshot(golden_dir, tmp_path):
    """csv -> mesh.npz -> csv on a committed reference snapshot: the values the reference's reader
    sees (pandas default parser) survive both directions unchanged."""
    import gzip
    import os
    import pandas as pd
    from mycelium_fea_project_b200 import snapshot_io
    src = os.path.join(golden_dir, "ref_results", "sim_20251117_181147")
    for f in ("nodes.csv", "elements.csv"):
        with gzip.open(os.path.join(src, f + ".gz")) as fi, open(tmp_path / f, "wb") as fo:
            fo.write(fi.read())
    nodes0 = pd.read_csv(tmp_path / "nodes.csv"); elems0 = pd.read_csv(tmp_path / "elements.csv")
    snapshot_io.csv_to_npz(str(tmp_path))
    z = np.load(tmp_path / "mesh.npz")
    assert np.array_equal(z["coords"], nodes0[["x", "y", "z"]].values)
    assert np.array_equal(z["n1"], elems0["n1"].values) and np.array_equal(z["n2"], elems0["n2"].values)
    os.remove(tmp_path / "nodes.csv"); os.remove(tmp_path / "elements.csv")
    snapshot_io.npz_to_csv(str(tmp_path))
    nodes1 = pd.read_csv(tmp_path / "nodes.csv"); elems1 = pd.read_csv(tmp_path / "elements.csv")
    assert list(nodes1.columns) == list(nodes0.columns) and list(elems1.columns) == list(elems0.columns)
    assert np.abs(nodes1[["x", "y", "z"]].values - nodes0[["x", "y", "z"]].values).max() <= 1e-15
    assert elems1.equals(elems0)


def test_stale_sidecar_is_ignored_and_node_ids_are_checked(tmp_path):
    """An edited CSV wins over an older mesh.npz; a node table whose node_id is not the row index is rejected
    (the reference addresses grips by node_id and elements by row position, src/fea_solver.py:82-83, 209-210)."""
    import os
    import pandas as pd
    import pytest
    from mycelium_fea_project_b200 import fea_solver as fs
    coords, n1, n2 = synth_network(12, 9, seed=2)
    d = str(tmp_path / "s")
    write_snapshot(d, coords, n1, n2, binary_sidecar=True)
    nodes = pd.read_csv(os.path.join(d, "nodes.csv"))
    nodes["x"] += 1.0
    nodes.to_csv(os.path.join(d, "nodes.csv"), index=False)
    t = os.path.getmtime(os.path.join(d, "mesh.npz"))
    os.utime(os.path.join(d, "nodes.csv"), (t + 5, t + 5))              # the CSV is now newer than the side-car
    c, _, _ = fs.load_snapshot(d)
    assert np.abs(c[:, 0] - (coords[:, 0] + 1.0)).max() <= 1e-12
    nodes["node_id"] = nodes["node_id"].values[::-1].copy()
    nodes.to_csv(os.path.join(d, "nodes.csv"), index=False)
    os.utime(os.path.join(d, "nodes.csv"), (t + 9, t + 9))
    with pytest.raises(ValueError, match="node_id"):
        fs.load_snapshot(d)


def _pandas_reference_writer(fea_dir, rec, n_elems):
    """What the reference writes (src/fea_solver.py:298-316): pandas DataFrames, one column per element / DOF."""
    import os
    import pandas as pd
    os.makedirs(fea_dir, exist_ok=True)
    cols = [f"elem_{i}" for i in range(n_elems)]
    steps = np.arange(1, len(rec["stress"]) + 1)
    for name, rows, columns in (("stress_record.csv", rec["stress"], cols), ("active_elements.csv", rec["active"], cols),
                                ("node_displacements.csv", rec["disp"], np.arange(len(rec["disp"][0])))):
        df = pd.DataFrame(rows, columns=columns)
        df["step"] = steps
        df.to_csv(os.path.join(fea_dir, name), index=False)
    pd.DataFrame(rec["force_disp"], columns=["total_displacement", "total_force"]).to_csv(
        os.path.join(fea_dir, "force_displacement.csv"), index=False)


CSVS = ("stress_record.csv", "active_elements.csv", "node_displacements.csv", "force_displacement.csv")


def test_result_writer_is_byte_identical_to_pandas(tmp_path):
    """fea_solver.write_results formats the wide CSVs itself (the DataFrame route was the ramp's wall clock at the
    reference's sizes); its bytes must be pandas' bytes -- shortest round-trip floats, exponents, -0.0, subnormals,
    True / False -- and non-finite values must take the pandas route (empty field)."""
    import filecmp
    from mycelium_fea_project_b200 import fea_solver as fs
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(3000) * 10.0 ** rng.integers(-320, 300, 3000),
                        [0.0, -0.0, 1.0, 2.0, 1e16, 1e15, 123456789012345678.0, 1e-5, 1e-4, 5e-324,
                         1.7976931348623157e308, -1e22, 1e21, 0.1, 1 / 3]])
    x = x[np.isfinite(x)]
    rec = {"stress": [x, -x, x[::-1].copy()], "active": [x > 0, x < 0, x == 0], "disp": [x[:60], x[60:120], x[120:180]],
           "force_disp": [[0.1, 0.2], [0.3, 1e-9], [-0.0, 5e-324]]}
    _pandas_reference_writer(str(tmp_path / "a"), rec, len(x))
    fs.write_results(str(tmp_path / "b"), rec, len(x))
    for f in CSVS:
        assert filecmp.cmp(tmp_path / "a" / f, tmp_path / "b" / f, shallow=False), f
    bad = {k: [np.array(r, copy=True) for r in v] if k != "force_disp" else v for k, v in rec.items()}
    bad["stress"][1][5] = np.nan
    bad["disp"][0][0] = np.inf
    _pandas_reference_writer(str(tmp_path / "c"), bad, len(x))
    fs.write_results(str(tmp_path / "d"), bad, len(x))
    for f in CSVS:
        assert filecmp.cmp(tmp_path / "c" / f, tmp_path / "d" / f, shallow=False), f


def test_result_writer_reproduces_the_reference_csv_bytes(golden_dir, tmp_path):
    """The ramp records of the CPU oracle on the reference's committed letter fixture, written by the PRODUCT's writer,
    are byte-equal to the CSVs the reference committed (results/test_I/fea_results)."""
    import filecmp
    import os
    import pandas as pd
    from oracle import fea_oracle as fo
    from mycelium_fea_project_b200 import fea_solver as fs
    src = os.path.join(golden_dir, "ref_results", "test_I")
    nodes = pd.read_csv(os.path.join(src, "nodes.csv"))
    elems = pd.read_csv(os.path.join(src, "elements.csv"))
    res = fo.fea_ramp(nodes[["x", "y", "z"]].values, elems["n1"].values, elems["n2"].values,
                      tol=0.5, disp_max=0.06, n_steps=40)
    rec = {"stress": res.stress, "active": res.active, "disp": res.disp, "force_disp": res.force_disp}
    fs.write_results(str(tmp_path / "fea_results"), rec, len(elems))
    for f in CSVS:
        assert filecmp.cmp(tmp_path / "fea_results" / f, os.path.join(src, "fea_results", f), shallow=False), f
