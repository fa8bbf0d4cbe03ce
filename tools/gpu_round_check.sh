# HISTORICAL: the exact command of the first round-1 GPU call of this session (K_e / assembly / fused-solve ncu captures).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
KREG='regex:^(ke_batch_kernel|edge_count_kernel|edge_emit_kernel|rs_hist_kernel|rs_scatter_kernel|block_count_kernel|row_ptr_kernel|fill_kernel)$'
timeout 240 python tools/ncu_ke_assembly.py --grid 2048 > gpurun_out/ke_asm_plain.log 2>&1
echo "rc=$?" >> gpurun_out/ke_asm_plain.log
MYC_NCU=1 timeout 300 ncu --set full --clock-control none -k "$KREG" -c 24 -f -o gpurun_out/prof_r1_ke_asm python tools/ncu_ke_assembly.py --grid 2048 > gpurun_out/ncu_ke_asm.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_ke_asm.log
timeout 420 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 90 python tools/ncu_fused_solve.py > gpurun_out/fused_plain.log 2>&1 && timeout 240 ncu --set full --clock-control none -k regex:pcg_fused -c 1 -f -o gpurun_out/prof_r1_fused_solve python tools/ncu_fused_solve.py > gpurun_out/ncu_fused.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_fused.log
timeout 300 python bench.py --steps 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
echo "rc=$?" >> gpurun_out/bench_n1.err
tail -3 gpurun_out/pytest_gpu.log; tail -c 600 gpurun_out/ke_asm_plain.log
