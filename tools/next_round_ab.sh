#!/bin/bash
# A/B matrix for the hooks that were written after round 1's GPU budget ended.
# Build the variants HERE (no GPU needed), then run the two blocks below under gpurun.
#
#   bash tools/next_round_ab.sh build
#   gpurun --timeout 500 -- 'bash tools/next_round_ab.sh gpu1'            # one GPU
#   gpurun --gpus 2 --timeout 400 -- 'bash tools/next_round_ab.sh gpu2'   # two GPUs
set -x
PKG=mycelium_fea_project_b200
case "$1" in
build)
  make -C $PKG/csrc -j8
  make -C $PKG/csrc -j8 TARGET=../libmycelium_fea_b200_bf.so  OBJDIR=../../build/obj_bf  EXTRA=-DMYC_BLOCK_FASTEST_WARPS
  make -C $PKG/csrc -j8 TARGET=../libmycelium_fea_b200_ldb.so OBJDIR=../../build/obj_ldb EXTRA=-DMYC_LIGHT_DIST_BARRIER
  ;;
gpu1)
  mkdir -p gpurun_out
  BF=$PWD/$PKG/libmycelium_fea_b200_bf.so
  # 1. block-fastest work split: whole suite, then same-box bench against the default
  MYC_LIB_PATH=$BF python -m pytest tests -m gpu -q > gpurun_out/ab_bf_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_bf_pytest.log
  python bench.py --steps 5 --no-cpu-baseline --no-hbm-roofline > gpurun_out/ab_default_bench.log 2>&1
  MYC_LIB_PATH=$BF python bench.py --steps 5 --no-cpu-baseline --no-hbm-roofline > gpurun_out/ab_bf_bench.log 2>&1
  # 2. short assembly sort: bit-equality test, then timing against the full sort
  MYC_TEST_SHORT_SORT=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "assembly" > gpurun_out/ab_shortsort_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_shortsort_pytest.log
  python tools/ncu_ke_assembly.py --grid 2048 > gpurun_out/ab_asm_full.log 2>&1
  MYC_ASM_SHORT_SORT=1 python tools/ncu_ke_assembly.py --grid 2048 > gpurun_out/ab_asm_short.log 2>&1
  tail -2 gpurun_out/ab_bf_pytest.log gpurun_out/ab_shortsort_pytest.log
  grep -h -o '"value": [0-9.]*' gpurun_out/ab_default_bench.log gpurun_out/ab_bf_bench.log
  grep -h -o '"assemble_ms": [0-9.]*' gpurun_out/ab_asm_full.log gpurun_out/ab_asm_short.log
  ;;
gpu2)
  mkdir -p gpurun_out
  LDB=$PWD/$PKG/libmycelium_fea_b200_ldb.so
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3"
  # 3. block6 on two GPUs (even-node cuts): parity test, then bench against block3
  MYC_TEST_DIST_BLOCK6=1 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/ab_dist_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_dist_pytest.log
  $RUN > gpurun_out/ab_n2_default.log 2>&1
  MYC_DIST_BLOCK6=1 $RUN > gpurun_out/ab_n2_block6.log 2>&1
  # 4. release/acquire form of the multi-GPU barrier: parity tests with the variant library, then bench
  MYC_LIB_PATH=$LDB python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/ab_ldb_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_ldb_pytest.log
  MYC_LIB_PATH=$LDB $RUN > gpurun_out/ab_n2_ldb.log 2>&1
  MYC_LIB_PATH=$LDB MYC_DIST_BLOCK6=1 $RUN > gpurun_out/ab_n2_ldb_block6.log 2>&1
  tail -2 gpurun_out/ab_dist_pytest.log gpurun_out/ab_ldb_pytest.log
  grep -h -o '"value": [0-9.]*' gpurun_out/ab_n2_default.log gpurun_out/ab_n2_block6.log gpurun_out/ab_n2_ldb.log gpurun_out/ab_n2_ldb_block6.log
  ;;
*) echo "usage: $0 build|gpu1|gpu2"; exit 2;;
esac
