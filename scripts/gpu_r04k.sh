#!/bin/bash
# r04k: stream probe incl. the reset pattern (two write streams)
OUT=gpurun_out/r04k; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe scripts/exp_stream_probe.cu && /tmp/probe | tee $OUT/stream_probe.log
