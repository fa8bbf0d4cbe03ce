"""B200-native FEA hot path of YiKwanwoo2/mycelium-fea-project.

Sub-modules:
  fea_solver  reference-compatible entry points (bar_stiffness_bulk, assemble_global_stiffness,
              solve_system, fea_solver) -- needs the built CUDA library and a B200
  device      device-resident objects over the C-ABI
  dist        row-partitioned multi-GPU solve (one process per GPU, NCCL)
  synth       synthetic grid-occupancy networks (pure numpy; benchmark inputs)
Importing the package itself is cheap; the CUDA library is bound on first use of
fea_solver / device / dist and raises ImportError if it has not been built.
"""
import importlib

__all__ = ["fea_solver", "device", "dist", "synth"]


def __getattr__(name):
    if name in __all__:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
