#!/bin/bash
# A/B of library builds on ONE box, back to back:  tools/ab_libs.sh <grid> <lib[,ENV=VAL...]> ...   (prints kernel_ms per run)
grid=$1; shift
for rep in 1 2; do
  for spec in "$@"; do
    lib=${spec%%,*}
    envs=""
    if [[ "$spec" == *,* ]]; then envs=$(echo "${spec#*,}" | tr ',' ' '); fi
    env $envs MYC_LIB_PATH=$PWD/$lib python tools/ncu_solve.py $grid amg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', 'rep$rep', 'grid', d['grid'], 'its', d['iterations'], 'kernel_ms %.2f' % d['kernel_ms'], 'GB/s %.0f' % d['GBs'], 'setup %.1f' % d['ms_setup'], 'relres %.2e' % d['relres'])"
  done
done
