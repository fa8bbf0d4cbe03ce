"""Import the UNMODIFIED reference solver from /root/reference -- TEST INFRASTRUCTURE ONLY.

/root/reference exists only in the build container, never on the GPU box, so
everything here degrades to ``None`` when the tree is absent and nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` depends on it at run time.

The reference module imports matplotlib at module top
(src/fea_solver_no_plotting.py:3,8-9) without using it on the no-plot path;
matplotlib is not installed here, so four stub modules are placed in
``sys.modules`` before the import.  No reference source is copied.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import shutil
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("MYC_REFERENCE_ROOT", "/root/reference")
_REF_FILE = os.path.join(REFERENCE_ROOT, "src", "fea_solver_no_plotting.py")
_cached = None


def available() -> bool:
    return os.path.isfile(_REF_FILE)


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    coll = types.ModuleType("matplotlib.collections")
    cols = types.ModuleType("matplotlib.colors")
    coll.LineCollection = object
    cols.Normalize = object
    mpl.pyplot, mpl.collections, mpl.colors = plt, coll, cols
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt,
                        "matplotlib.collections": coll, "matplotlib.colors": cols})


def load_reference():
    """The reference module object, or None when /root/reference is absent."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    _stub_matplotlib()
    spec = importlib.util.spec_from_file_location("ref_fea_solver_no_plotting", _REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod


@contextlib.contextmanager
def patched_constants(mod, **consts):
    """Temporarily override module-level constants (N_STEPS, DISPLACEMENT_MAX, ...),
    which the reference reads at call time."""
    old = {k: getattr(mod, k) for k in consts}
    try:
        for k, v in consts.items():
            setattr(mod, k, v)
        yield mod
    finally:
        for k, v in old.items():
            setattr(mod, k, v)


def run_reference(results_dir_src, tol, **consts):
    """Run the reference's fea_solver() on a copy of ``results_dir_src``'s two input
    CSVs (the reference tree is read-only) and return the temp dir holding
    ``fea_results/``.  Caller removes it."""
    mod = load_reference()
    if mod is None:
        raise RuntimeError("reference tree not present")
    tmp = tempfile.mkdtemp(prefix="myc_ref_")
    for f in ("nodes.csv", "elements.csv"):
        shutil.copy(os.path.join(results_dir_src, f), os.path.join(tmp, f))
    with patched_constants(mod, **consts), open(os.devnull, "w") as devnull, \
            contextlib.redirect_stdout(devnull):
        mod.fea_solver(tmp, tol=tol)
    return tmp
