#!/bin/bash
# Kernel-tuning experiment: build libpomdp_b200 variants (CTA size / min-blocks) and time bench.py with each.
# Usage on the GPU box: bash scripts/exp_variants.sh <tag> "512 2" "1024 1" "256 4" ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT /tmp/variants
for cfg in "$@"; do
  set -- $cfg
  so=/tmp/variants/lib_$1_$2.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
       -DPOMDP_STEP_THREADS=$1 -DPOMDP_STEP_MINB=$2 ${EXTRA_NVCC} -o $so gym_pomdp_b200/csrc/pomdp_kernels.cu 2>&1 | grep -E "error" 
  echo "== variant threads=$1 minb=$2"
  POMDP_B200_LIB=$so python bench.py --no-cpu --steps 1000 --e2e-steps 2 ${BENCH_ARGS} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   value %.4e  us/launch %.2f  frac %.3f' % (d['value'], d['ms_per_step']*1e3, d['roofline']['frac']))
    else: print(l.rstrip()[:200])
" | tee -a $OUT/variants.log
done
