#!/bin/bash
# r04o: fused histogram + all-reduce over peer memory.  1 GPU: the kernel path with a hand-made peer table;
# N > 1 GPUs: bench's collective row (NCCL vs fused, equality + timing)
N=${1:-1}; OUT=gpurun_out/r04o; mkdir -p $OUT
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests/test_edge_cases.py tests/test_fullsize_parity.py -m gpu -q -x -k "hist" 2>&1 | tail -3 | tee $OUT/pytest.log
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 200 --warmup 10 --no-cpu 2> $OUT/bench_${N}.err > $OUT/bench_${N}.json
  tail -5 $OUT/bench_${N}.err
  python - <<PY
import json
d = json.loads(open("$OUT/bench_${N}.json").read().strip().splitlines()[-1])
print(d["value"], d["collective"])
PY
fi
