#!/bin/bash
# r04p: fused histogram + all-reduce vs NCCL, eager and graph-captured, N GPUs
N=${1:-8}; OUT=gpurun_out/r04p; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    scripts/bench_fused_hist.py --out $OUT/fused_hist_${N}gpu.json 2> $OUT/err_$N.log | tail -2
tail -3 $OUT/err_$N.log
