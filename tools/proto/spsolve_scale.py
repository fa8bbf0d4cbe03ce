"""How far does the reference's direct solve (oracle.solve_system -> SuperLU) go on this host?"""
import sys, time, resource
sys.path.insert(0, ".")
import numpy as np
from oracle import fea_oracle as fo
from mycelium_fea_project_b200.synth import synth_network
for N in [int(a) for a in sys.argv[1:]]:
    coords, n1, n2 = synth_network(N)
    t0 = time.perf_counter()
    K = fo.assemble_global_stiffness(coords, n1, n2, np.ones(len(n1), bool))
    t1 = time.perf_counter()
    hi, lo = fo.grip_nodes(coords, 1.5, 1)
    kd, kv = fo.build_bc(hi, lo, 0.02, -0.02, 1)
    U = fo.solve_system(K, kd, kv)
    t2 = time.perf_counter()
    print(N, "n_dof", K.shape[0], "asm_vec %.1fs solve %.1fs maxrss %.1f GB" % (t1 - t0, t2 - t1, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6), flush=True)
