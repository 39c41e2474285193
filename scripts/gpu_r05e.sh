#!/bin/bash
# r05e: belief histogram loop variants (loads in flight per trip x software prefetch of the next trip), built on the CPU box
# into build_variants/ and selected with POMDP_B200_LIB.   gpurun -- bash scripts/gpu_r05e.sh
OUT=gpurun_out/r05e; mkdir -p $OUT
for v in 4_0 4_1 2_1 2_0 8_0; do
  echo "== variant inflight_prefetch=$v" | tee -a $OUT/hist_variants.log
  POMDP_B200_LIB=$PWD/build_variants/lib_hist_$v.so timeout 300 python scripts/bench_configs.py --kernels belief_hist --steps 200 2>&1 \
    | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   %-50s %7.2f us  %5.1f %%' % (d['config'], d['us_per_launch'], 100 * d['frac_of_peak']))
" | tee -a $OUT/hist_variants.log
done
