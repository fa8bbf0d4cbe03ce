"""ctypes binding of libmycelium_fea_b200.so (the C-ABI declared in include/mycelium_fea.h).

There is no CPU fallback: importing this module without the built library raises, and
creating a context without a B200 raises.  Build with ``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C mycelium_fea_project_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MYC_LIB_PATH") or os.path.join(_HERE, "libmycelium_fea_b200.so")   # override: A/B builds

MYC_OK = 0
MYC_ERR_BAD_ARG = -1
MYC_ERR_CUDA = -2
MYC_ERR_NCCL = -3
MYC_ERR_NOT_CONVERGED = -4
MYC_ERR_BREAKDOWN = -5
MYC_ERR_CAPACITY = -6
MYC_ERR_STATE = -7
MYC_PC_JACOBI = 0
MYC_PC_BLOCK3 = 1
MYC_PC_BLOCK6 = 2
MYC_PC_BLOCK12 = 3
MYC_PC_AMG = 4
PRECONDITIONERS = {"jacobi": MYC_PC_JACOBI, "block3": MYC_PC_BLOCK3, "block6": MYC_PC_BLOCK6, "block12": MYC_PC_BLOCK12,
                   "amg": MYC_PC_AMG}

_ERR_NAMES = {-1: "BAD_ARG", -2: "CUDA", -3: "NCCL", -4: "NOT_CONVERGED", -5: "BREAKDOWN",
              -6: "CAPACITY", -7: "STATE"}


class MyceliumFeaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"MYC_ERR_{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class NotConverged(MyceliumFeaError):
    pass


if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA extension is the product and there is no CPU path. "
        "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i64 = C.c_int64
_f64 = C.c_double
_int = C.c_int
_pi64 = C.POINTER(C.c_int64)
_pf64 = C.POINTER(C.c_double)

# name -> argtypes; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "myc_abi_version": [],
    "myc_create": [_int, C.POINTER(_p)],
    "myc_destroy": [_p],
    "myc_last_error": [_p],
    "myc_launch_count": [_p],
    "myc_set_csr_hint": [_p, _int],
    "myc_csr_is_block3": [_p, _i64, _p, _p, C.POINTER(C.c_int), _p],
    "myc_profile_reset": [_p, _int],
    "myc_profile_get": [_p, _pf64],
    "myc_bar_stiffness_bulk": [_p, _p, _p, _i64, _f64, _f64, _f64, _p, _p, _p],
    "myc_assemble_symbolic": [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _pi64, _p],
    "myc_assemble_numeric": [_p, _p, _p, _p, _f64, _f64, _f64, _i64, _p, _p, _p, _p],
    "myc_apply_dirichlet": [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _i64, _f64, _p, _p, _p, _p],
    "myc_block3_inverse": [_p, _i64, _i64, _p, _p, _p, _p, _f64, _p, _p],
    "myc_block_inverse_size": [_int, _i64],
    "myc_block_inverse_packed": [_p, _int, _i64, _i64, _p, _p, _p, _p, _f64, _p, _p],
    "myc_amg_setup": [_p, _i64, _i64, _i64, _p, _p, _p, _p, _f64, C.POINTER(C.c_int), _p],
    "myc_amg_level_info": [_p, _int, _pi64, _p, _p],
    "myc_reduce_csr": [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _pi64, _pi64, _p],
    "myc_spmv": [_p, _i64, _p, _p, _p, _p, _p, _p],
    "myc_pcg_solve": [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _int, _f64, _f64, _f64, _i64, _p,
                      _pi64, _pf64, _p],
    "myc_true_residual": [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _f64, _p, _pf64, _p],
    "myc_merge_solution": [_p, _i64, _i64, _p, _p, _p, _p, _p],
    "myc_gather_sum": [_p, _p, _p, _i64, _pf64, _p],
    "myc_strain_update": [_p, _p, _p, _p, _i64, _p, _f64, _f64, _p, _p, _pi64, _p],
    "myc_dist_unique_id": [C.c_char_p, _p],
    "myc_dist_init": [_p, C.c_char_p, _p, _int, _int],
    "myc_dist_set_plan": [_p, _p, _p, _p, _p, _p],
    "myc_dist_peer_alloc": [_p, _i64, _p],
    "myc_dist_peer_open": [_p, _p],
    "myc_dist_peer_disable": [_p],
    "myc_dist_release_peers": [_p],
    "myc_halo_exchange": [_p, _p, _p],
    "myc_allreduce_sum": [_p, _pf64, _int, _p],
    "myc_allgather_owned": [_p, _p, _p],
    "myc_load_case_host": [_p, _p, _p, _p, _p, _i64, _i64, _f64, _f64, _f64, _p, _p, _i64, _f64, _int,
                           _f64, _i64, _p, _i64, _p, _pf64, _pi64, _pf64, _pi64, _pf64, _pf64],
}
_RESTYPES = {"myc_last_error": C.c_char_p, "myc_launch_count": C.c_int64, "myc_block_inverse_size": C.c_int64}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, C.c_int)

if lib.myc_abi_version() != 1:
    raise ImportError("libmycelium_fea_b200.so ABI version mismatch")


def nccl_library_path():
    """Path of the torch-bundled libnccl.so.2 (what torch.distributed itself uses)."""
    try:
        import nvidia.nccl  # type: ignore
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.isfile(cand):
                return cand
    except Exception:
        pass
    return "libnccl.so.2"


def check(ctx_handle, rc, allow_not_converged=False):
    if rc == MYC_OK:
        return rc
    msg = lib.myc_last_error(ctx_handle)
    msg = msg.decode() if msg else ""
    if rc == MYC_ERR_NOT_CONVERGED:
        if allow_not_converged:
            return rc
        raise NotConverged(rc, msg)
    raise MyceliumFeaError(rc, msg)
