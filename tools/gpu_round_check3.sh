# HISTORICAL: the exact command of one round-1 GPU call (results in profiles/r1_block_jacobi_groups.md).  The variant
# libraries it names (_lb, _b6full) were A/B builds whose winners are now the defaults; see tools/next_round_ab.sh.
# Final round-1 check of the shipped defaults (block6 + release/acquire barrier + staged assembly fill),
# A/B of the full-row block6 layout, refreshed ncu launch list and fused-kernel capture.
set -x
mkdir -p gpurun_out
FULL=$PWD/mycelium_fea_project_b200/libmycelium_fea_b200_b6full.so
timeout 200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_final.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_final.log
timeout 200 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "rc=$?" >> gpurun_out/bench_final.err
MYC_LIB_PATH=$FULL timeout 100 python bench.py --steps 3 --no-cpu-baseline --no-hbm-roofline > gpurun_out/bench_b6full.log 2> gpurun_out/bench_b6full.err
MYC_LIB_PATH=$FULL timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "solve or ramp or host or pcg or block" > gpurun_out/pytest_b6full.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_b6full.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_block6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-hbm-roofline > gpurun_out/ncu_launches.log 2>&1
timeout 60 python tools/ncu_fused_solve.py > gpurun_out/fused_plain_block6.log 2>&1 && timeout 150 ncu --set full --clock-control none -k regex:pcg_fused -c 1 -f -o gpurun_out/prof_r1_fused_solve_block6 python tools/ncu_fused_solve.py > gpurun_out/ncu_fused_block6.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_final.log gpurun_out/pytest_b6full.log
grep -h -o '"value": [0-9.]*' gpurun_out/bench_final.log gpurun_out/bench_b6full.log
