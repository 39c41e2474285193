#!/bin/bash
# r04f: (1) compute-free probe of the step kernels' six-stream traffic pattern = the roof for this read:write mix and the
# fixed cost of a launch at the small BASELINE sizes; (2) CTA-shape variants of the Network step after the joint draw
OUT=gpurun_out/r04f; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe scripts/exp_stream_probe.cu && /tmp/probe | tee $OUT/stream_probe.log
ONLY="Network" bash scripts/exp_network_variants.sh r04f "" "-DPOMDP_STEP_THREADS=256 -DPOMDP_STEP_MINB=4" "-DPOMDP_STEP_THREADS=256 -DPOMDP_STEP_MINB=5" "-DPOMDP_STEP_THREADS=384 -DPOMDP_STEP_MINB=3" "-DPOMDP_STEP_THREADS=1024 -DPOMDP_STEP_MINB=1" "-DPOMDP_STEP_THREADS=128 -DPOMDP_STEP_MINB=8"
