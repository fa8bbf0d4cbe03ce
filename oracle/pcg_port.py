"""ctypes wrapper of oracle/pcg_port.c -- TEST INFRASTRUCTURE ONLY (multi-threaded CPU baseline that
restates the reference's PETSc path: MatZeroRowsColumns + 1e-12 shift + KSPCG/PCJACOBI,
src/fea_petsc.cpp:303-341).  Built by `make -C oracle` (called from __graft_entry__.build())."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libpcg_port.so")


def available() -> bool:
    return os.path.isfile(_LIB)


def _lib():
    lib = C.CDLL(_LIB)
    lib.pcg_port_threads.restype = C.c_int
    lib.pcg_port_solve.restype = C.c_int64
    lib.pcg_port_solve.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_int64, C.c_void_p, C.POINTER(C.c_double)]
    lib.pcg_port_zero_rows_cols.restype = None
    lib.pcg_port_zero_rows_cols.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_double, C.c_void_p, C.c_void_p]
    return lib


def threads() -> int:
    return int(_lib().pcg_port_threads())


def solve_system_petsc_style(K, known_dofs, known_vals, rtol=1e-10, max_iters=10_000_000, reg=1e-12):
    """U, iterations, relres for one load step the way the reference's PETSc path computes it.
    ``K`` is the scipy CSR from assemble_global_stiffness (int32 indices)."""
    lib = _lib()
    n = K.shape[0]
    rp = np.ascontiguousarray(K.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(K.indices, dtype=np.int32)
    val = np.array(K.data, dtype=np.float64, copy=True)
    is_known = np.zeros(n, dtype=np.uint8)
    x_known = np.zeros(n)
    is_known[known_dofs] = 1
    x_known[known_dofs] = known_vals
    b = np.empty(n)
    extra = np.empty(n)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.pcg_port_zero_rows_cols(n, p(rp), p(ci), p(val), p(is_known), p(x_known), float(reg), p(b), p(extra))
    x = np.empty(n)
    rel = C.c_double(0.0)
    it = lib.pcg_port_solve(n, p(rp), p(ci), p(val), p(extra), p(b), float(rtol), int(max_iters), p(x), C.byref(rel))
    return x, int(it), float(rel.value)
