"""CPU checks of the boundary: the library loads without a GPU, exports every symbol the header
declares, and the host-side logic (grips, BC sets) matches the oracle bit for bit."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mycelium_fea.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(myc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mycelium_fea_project_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in include/mycelium_fea.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert _lib.lib.myc_abi_version() == 1


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mycelium_fea_project_b200 import _lib
    h = ctypes.c_void_p()
    rc = _lib.lib.myc_create(0, ctypes.byref(h))
    assert rc == _lib.MYC_ERR_CUDA and not h.value
    assert b"no CPU path" in _lib.lib.myc_last_error(None)
    from mycelium_fea_project_b200 import fea_solver as fs
    with pytest.raises(RuntimeError):
        fs.bar_stiffness_bulk(np.zeros((1, 3)), np.ones((1, 3)))


def test_constants_match_oracle():
    from mycelium_fea_project_b200 import fea_solver as fs
    from oracle import fea_oracle as fo
    for k in ("E_mod", "A", "I", "N_STEPS", "DISPLACEMENT_MAX", "MAX_STRAIN", "MAX_STRESS", "GRIP_LENGTH",
              "REGULARISATION"):
        assert getattr(fs, k) == getattr(fo, k), k


@pytest.mark.parametrize("N,tol,axis,comp", [(64, 0.5, 1, 1), (64, 1.5, 1, 1), (128, 1.5, 0, 0), (64, 2.0, 1, 0)])
def test_bc_sets_bit_exact(N, tol, axis, comp):
    from mycelium_fea_project_b200 import fea_solver as fs
    from mycelium_fea_project_b200.synth import synth_network
    from oracle import fea_oracle as fo
    coords, _, _ = synth_network(N)
    hi, lo = fs.grip_nodes(coords, tol, axis)
    ho, loo = fo.grip_nodes(coords, tol, axis)
    assert np.array_equal(hi, ho) and np.array_equal(lo, loo)
    kd, kv = fs.build_bc(hi, lo, 0.0123, -0.0123, comp)
    kdo, kvo = fo.build_bc(ho, loo, 0.0123, -0.0123, comp)
    assert kd.dtype == kdo.dtype and np.array_equal(kd, kdo)      # ordered list, not just the set
    assert np.array_equal(kv, kvo)


def test_bc_sets_golden(golden_dir):
    from mycelium_fea_project_b200 import fea_solver as fs
    from mycelium_fea_project_b200.synth import synth_network
    g = np.load(os.path.join(golden_dir, "solve_synth128.npz"))
    coords, _, _ = synth_network(128)
    hi, lo = fs.grip_nodes(coords, float(g["tol"]))
    kd, kv = fs.build_bc(hi, lo, 0.02, -0.02)
    assert np.array_equal(kd, g["known_dofs"]) and np.array_equal(kv, g["known_vals"])
