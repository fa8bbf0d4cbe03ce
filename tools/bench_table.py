"""Markdown table of the committed bench lines:  python tools/bench_table.py profiles/r2_bench_n{1,2,4,8}.json"""
import json
import sys


def load(p):
    return json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1])


rows = [load(p) for p in sys.argv[1:]]
v1 = rows[0]["value"] / rows[0]["n_gpus"]
hdr = ["N", "DOF (X / Y)", "value MDOF/s", "e2e MDOF/s", "vs N x (N=1)", "ms/step", "assemble / setup / solve ms", "its X / Y",
       "roofline.frac", "true relres X / Y"]
print("| " + " | ".join(hdr) + " |")
print("|" + "---|" * len(hdr))
for d in rows:
    lc = d["load_cases"]
    print(f"| {d['n_gpus']} | {d['config']['n_dof']['X'] / 1e6:.1f} M / {d['config']['n_dof']['Y'] / 1e6:.1f} M | {d['value']:.1f} | "
          f"{d['e2e']['value']:.1f} | {d['value'] / (v1 * d['n_gpus']):.2f} | {d['ms_per_step']:.1f} | "
          f"{d['ms_assemble']:.1f} / {d['ms_setup']:.1f} / {d['ms_solve']:.1f} | {lc['X']['iterations']} / {lc['Y']['iterations']} | "
          f"{d['roofline']['frac']:.2f} | {lc['X']['true_relres']:.1e} / {lc['Y']['true_relres']:.1e} |")
print()
hdr = ["N", "strong_4096: its", "assemble / setup / solve ms", "total ms", "speed-up", "us / iteration", "total_force",
       "parity: relL2(U) vs 1 GPU", "force rel. diff"]
print("| " + " | ".join(hdr) + " |")
print("|" + "---|" * len(hdr))
t1 = None
for d in rows:
    s = d.get("strong_4096")
    if not s or "error" in s:
        continue
    tot = s["ms_assemble"] + s["ms_setup"] + s["ms_solve"]
    t1 = t1 or tot
    par = s.get("parity", {})
    print(f"| {d['n_gpus']} | {s['iterations']} | {s['ms_assemble']:.1f} / {s['ms_setup']:.1f} / {s['ms_solve']:.1f} | {tot:.1f} | "
          f"{t1 / tot:.2f} | {s['us_per_iteration']:.0f} | {s['total_force']:.12e} | "
          f"{par.get('relL2_U_vs_1gpu', float('nan')):.1e} | {par.get('force_rel_diff', float('nan')):.1e} |")
for d in rows:
    c = d.get("config4_8192")
    if c and "error" not in c:
        print()
        for k, s in c.items():
            print(f"config4_8192 {k}: {s['n_dof'] / 1e6:.1f} M DOF, {s['iterations']} iterations, assemble / setup / solve "
                  f"{s['ms_assemble']:.1f} / {s['ms_setup']:.1f} / {s['ms_solve']:.1f} ms, true relres {s['true_relres']:.1e}, "
                  f"{s['mdof_per_s']:.0f} MDOF/s")
