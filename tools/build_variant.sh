#!/bin/bash
# A/B build of the multigrid solver kernel only:  tools/build_variant.sh <name> [nvcc -D flags...]
# -> build/libmyc_<name>.so = the shipped objects (build/obj, built by `make`) with pcg_amg.o recompiled under the flags.
# Select it at run time with MYC_LIB_PATH (tools/ab_libs.sh).  amg_sweep.cuh is included by pcg_amg.cu only.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build/ab_$name
nvcc "$@" -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v \
     -c $root/mycelium_fea_project_b200/csrc/pcg_amg.cu -o $root/build/ab_$name/pcg_amg.o 2> $root/build/ab_$name/ptxas.log
objs=$(ls $root/build/obj/*.o | grep -v pcg_amg.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $root/build/libmyc_$name.so $objs $root/build/ab_$name/pcg_amg.o -lcudart -ldl
grep -A2 "pcg_amg_kernelILb0ELb1" $root/build/ab_$name/ptxas.log | grep -E "spill|Used" | tr '\n' ' '; echo " -> build/libmyc_$name.so"
