#!/bin/bash
# Network step tuning: build libpomdp_b200 variants and time the Network-v0 step with each (scripts/bench_configs.py).
# Usage on the GPU box: bash scripts/exp_network_variants.sh <tag> "<nvcc -D flags>" ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT /tmp/variants
i=0
for flags in "$@"; do
  so=/tmp/variants/lib_net_$i.so; i=$((i+1))
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $flags -o $so gym_pomdp_b200/csrc/pomdp_kernels.cu 2>&1 | grep -E "error"
  echo "== variant [$flags]" | tee -a $OUT/variants.log
  POMDP_B200_LIB=$so python scripts/bench_configs.py --no-rollout --only "${ONLY:-Network}" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        if d['kernel'] in ('step', 'step_packed', 'reset'): print('   %-34s %-12s %8.2f us  %.3f of peak' % (d['config'], d['kernel'], d['us_per_launch'], d['frac_of_peak']))
" | tee -a $OUT/variants.log
done
