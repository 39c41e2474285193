#!/bin/bash
TAG=${1:-r03e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvcc -O3 -arch=sm_100a -o /tmp/exp_host_pipe2 scripts/exp_host_pipe2.cu && /tmp/exp_host_pipe2 | tee $OUT/host_pipe2.log
echo "== launch-shape variants: Network + Tag step"
ONLY="Network-v0" bash scripts/exp_network_variants.sh $TAG "" "-DPOMDP_STEP_THREADS=512 -DPOMDP_STEP_MINB=3" "-DPOMDP_STEP_THREADS=256 -DPOMDP_STEP_MINB=6" "-DPOMDP_STEP_THREADS=384 -DPOMDP_STEP_MINB=4" "-DPOMDP_STEP_THREADS=128 -DPOMDP_STEP_MINB=12" "-DPOMDP_STEP_THREADS=256 -DPOMDP_STEP_MINB=5"
ONLY="Tag-v0 B=2^22" bash scripts/exp_network_variants.sh $TAG "" "-DPOMDP_STEP_THREADS=512 -DPOMDP_STEP_MINB=3" "-DPOMDP_STEP_THREADS=256 -DPOMDP_STEP_MINB=6" "-DPOMDP_STEP_THREADS=384 -DPOMDP_STEP_MINB=4"
