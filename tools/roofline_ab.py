"""A/B helper: the bench's `roofline_hbm` (CSR SpMV with L2 flush + assembly) and, with --block6, its
`roofline_block6` (block-Jacobi persistent kernel, 2048^2 Y load case) for the library MYC_LIB_PATH selects.

    MYC_LIB_PATH=build/libmyc_x.so python tools/roofline_ab.py [--block6]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mycelium_fea_project_b200 import device as dv, fea_solver as fs
from mycelium_fea_project_b200._lib import lib

ap = argparse.ArgumentParser()
ap.add_argument("--block6", action="store_true")
ap.add_argument("--grid", type=int, default=2048)
args = ap.parse_args()
ctx = dv.Context.get()
peak, src = 6555.8, "fixed for the A/B"
h = bench.hbm_roofline(ctx, dv, fs, peak, src, N=args.grid)
out = {"lib": os.environ.get("MYC_LIB_PATH", "default"), "carveout": os.environ.get("MYC_CARVEOUT"),
       "spmv_us": round(h["avg_launch_us"], 1), "spmv_GBs": round(h["achieved"]), "asm_ms": round(h["assembly"]["ms"], 3)}
if args.block6:
    b = bench.block6_roofline(ctx, dv, fs, lib, args, peak, src)
    out.update({"block6_its": b["iterations"], "block6_us_per_it": round(b["us_per_iteration"], 1), "block6_GBs": round(b["achieved"])})
print(json.dumps(out))
