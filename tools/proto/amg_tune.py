"""CPU parameter study of the multigrid preconditioner (research record, not product code).

Runs oracle/amg_oracle.amg_pcg (the numpy restatement of csrc/amg_setup.cu + csrc/pcg_amg.cu) on a synthetic grid
with the module's constants overridden, and prints iterations, level sizes and a cost estimate per iteration from
the per-phase times measured on a B200 at 2048^2 (profiles/r2_amg_phase_timing.md): a level costs its four phases in
proportion to its size, but never less than the latency floor of ~9 us per phase.

    python tools/proto/amg_tune.py N [--case Y] name=value[,name=value...] ...
e.g. python tools/proto/amg_tune.py 512 MIN_NODES=200 MIN_NODES=2000,COARSE_SWEEPS=16
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import amg_oracle as ao
from oracle import fea_oracle as fo
from mycelium_fea_project_b200.synth import synth_network

# measured us per iteration at 2048^2 (2,796,332 nodes on level 0): D0, residual sweep, prolongation, post-smoothing
L0_NODES = 2796332.0
L0_PHASE_US = (223.0, 182.0, 55.0, 227.0)
CG_US = 176.0
FLOOR_US = 9.0


def problem(N, case):
    coords, n1, n2 = synth_network(N)
    K = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
    axis, comp = {"Y": (1, 1), "X": (0, 0)}[case]
    hi, lo = fo.grip_nodes(coords, 1.5, axis)
    kd, kv = fo.build_bc(hi, lo, 0.02, -0.02, comp)
    n = K.shape[0]
    free = np.ones(n, bool)
    free[kd] = False
    u = np.zeros(n)
    u[kd] = kv
    b = -(K @ u)
    b[~free] = 0.0
    return K, free, b


def cost_us(levels, scale_nodes):
    """estimated us per iteration if the level sizes are scaled so that level 0 has L0_NODES nodes"""
    total = CG_US
    for l, L in enumerate(levels):
        f = L.n * scale_nodes / L0_NODES
        if l == len(levels) - 1 and l > 0:
            total += max(FLOOR_US, L0_PHASE_US[0] * f) + (ao.COARSE_SWEEPS - 1) * max(0.45 * FLOOR_US, L0_PHASE_US[3] * f)
        else:
            total += sum(max(FLOOR_US, p * f) for p in L0_PHASE_US)
    return total


def main():
    N = int(sys.argv[1])
    args = sys.argv[2:]
    case = "Y"
    if args and args[0] == "--case":
        case = args[1]
        args = args[2:]
    K, free, b = problem(N, case)
    defaults = {k: getattr(ao, k) for k in ("OMEGA", "SCALE", "COARSE_SWEEPS", "MIN_NODES", "MAX_RATIO", "PC_FP32")}
    for spec in args or ["MIN_NODES=200"]:
        for k, v in defaults.items():
            setattr(ao, k, v)
        for kv in spec.split(","):
            k, v = kv.split("=")
            setattr(ao, k, type(defaults[k])(float(v)) if not isinstance(defaults[k], bool) else v == "1")
        t0 = time.time()
        x, its, levels = ao.amg_pcg(K, free, b, rtol=1e-10, maxit=2000)
        n0 = levels[0].n
        c = cost_us(levels, L0_NODES / n0)
        print(f"{spec:50s} its {its:4d}  levels {len(levels):2d} {[l.n for l in levels]}  est {c:7.0f} us/it  "
              f"-> {its * c / 1e3:7.1f} ms (scaled to 2048^2 sizes, same its)  [{time.time() - t0:.0f} s]", flush=True)


if __name__ == "__main__":
    main()
