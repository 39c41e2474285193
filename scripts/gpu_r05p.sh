#!/bin/bash
# r05p: N-GPU scaling bench (weak scaling, index shards, e2e copy ceiling with all ranks at once, config-5 collective row)
N=${1:-8}; OUT=gpurun_out/r05p; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_$N.csv
nvidia-smi topo -m > $OUT/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 1000 --warmup 10 --no-cpu 2> $OUT/bench_${N}.err | tee $OUT/bench_${N}.json | cut -c1-400
tail -3 $OUT/bench_${N}.err
echo "== reference arm x$N (driver-style)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 2>> $OUT/bench_${N}.err | tee $OUT/bench_ref_${N}.json | cut -c1-300
python - <<PY
import json
d = json.loads(open("$OUT/bench_${N}.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_ceiling"))
print(json.dumps(d["collective"])[:1500])
PY
