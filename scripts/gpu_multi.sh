#!/bin/bash
# N-GPU check of the scaling bench: gpurun --gpus N -- bash scripts/gpu_multi.sh <tag> N
TAG=${1:-multi}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.csv
echo "== bench x$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 1000 --warmup 10 --no-cpu 2> $OUT/bench_${N}.err | tee $OUT/bench_${N}.json
tail -3 $OUT/bench_${N}.err
echo "== reference arm x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 2>> $OUT/bench_${N}.err | tee $OUT/bench_ref_${N}.json
echo "== belief histogram all-reduce over NCCL (Rock(15,15), 2^25 global)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    scripts/nccl_hist.py 2>> $OUT/bench_${N}.err | tee $OUT/nccl_hist_${N}.json
tail -3 $OUT/bench_${N}.err
