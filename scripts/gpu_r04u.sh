#!/bin/bash
# r04u: N-GPU scaling bench (weak scaling, index shards, e2e copy ceiling with all ranks at once, config-5 collective row)
N=${1:-8}; OUT=gpurun_out/r04u; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_$N.csv
nvidia-smi topo -m > $OUT/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 1000 --warmup 10 --no-cpu 2> $OUT/bench_${N}.err | tee $OUT/bench_${N}.json | cut -c1-400
tail -3 $OUT/bench_${N}.err
