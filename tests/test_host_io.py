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
