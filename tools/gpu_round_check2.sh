# HISTORICAL: the exact command of one round-1 GPU call (results in profiles/r1_block_jacobi_groups.md).  The variant
# libraries it names (_lb, _b6full) were A/B builds whose winners are now the defaults; see tools/next_round_ab.sh.
# Round-1 validation of: staged assembly fill, block6/block12 preconditioners, light-barrier build.
set -x
mkdir -p gpurun_out
LB=$PWD/mycelium_fea_project_b200/libmycelium_fea_b200_lb.so
# A: whole GPU suite with the shipped defaults
timeout 300 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_A.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_A.log
# headline candidates, same box
for pc in block3 block6 block12; do
  timeout 200 python bench.py --steps 3 --precond $pc --no-cpu-baseline --no-hbm-roofline > gpurun_out/bench_$pc.log 2> gpurun_out/bench_$pc.err; echo "rc=$?" >> gpurun_out/bench_$pc.err
done
MYC_LIB_PATH=$LB timeout 200 python bench.py --steps 3 --precond block3 --no-cpu-baseline --no-hbm-roofline > gpurun_out/bench_block3_lb.log 2> gpurun_out/bench_block3_lb.err
MYC_LIB_PATH=$LB timeout 200 python bench.py --steps 3 --precond block12 --no-cpu-baseline --no-hbm-roofline > gpurun_out/bench_block12_lb.log 2> gpurun_out/bench_block12_lb.err
# B: whole suite with block12 as the default preconditioner; C: with the light-barrier build (+ block12)
MYC_PCG_PRECOND=block12 timeout 300 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_B.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_B.log
MYC_LIB_PATH=$LB MYC_PCG_PRECOND=block12 timeout 300 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_C.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_C.log
MYC_PCG_PRECOND=block6 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "solve or ramp or host or other_load" > gpurun_out/pytest_D.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_D.log
MYC_LIB_PATH=$LB timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "solve or ramp or host or pcg" > gpurun_out/pytest_E.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_E.log
# assembly: staged vs direct fill, then one ncu capture of the staged kernel
timeout 120 python tools/ncu_ke_assembly.py --grid 2048 > gpurun_out/ke_asm_staged.log 2>&1
MYC_ASM_DIRECT_FILL=1 timeout 120 python tools/ncu_ke_assembly.py --grid 2048 > gpurun_out/ke_asm_direct.log 2>&1
MYC_NCU=1 timeout 200 ncu --set full --clock-control none -k regex:fill_staged_kernel -c 1 -f -o gpurun_out/prof_r1_fill_staged python tools/ncu_ke_assembly.py --grid 2048 > gpurun_out/ncu_fill_staged.log 2>&1
# HBM-bound size: full 2048^2 Y solves
for pc in block12 block6; do
  timeout 120 python tools/perf_probe.py --grids 2048 --precond $pc > gpurun_out/probe2048_$pc.log 2>&1
done
tail -2 gpurun_out/pytest_A.log gpurun_out/pytest_B.log gpurun_out/pytest_C.log gpurun_out/pytest_D.log gpurun_out/pytest_E.log
grep -h -o '"value": [0-9.]*' gpurun_out/bench_*.log
