#!/bin/bash
# r03g (2 GPUs): GPU parity suite on one of them, then the scaling bench line under torchrun (configs + collective rows, e2e ceiling)
TAG=${1:-r03g}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.csv
nvidia-smi topo -m > $OUT/topo.txt 2>&1
if [ "$N" = "2" ]; then echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $OUT/pytest_gpu.log; fi
echo "== bench x$N (driver's flags)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 2> $OUT/bench_${N}.err | tee $OUT/bench_${N}.json | cut -c1-300
tail -3 $OUT/bench_${N}.err
python - "$OUT/bench_${N}.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d["e2e"]
print("value", d["value"], "ms/step", d["ms_per_step"], "timing", d["timing"]["mode"][:40])
print("e2e", e["value"], "ms", e["ms_per_step"], "ceiling_ms", e["ceiling_ms"], "frac", e["frac_of_ceiling"], "zero_copy", (e.get("zero_copy") or {}).get("ms_per_step"))
print("collective", d["collective"])
for r in d["configs"] or []:
    print("   %-48s %-6s %8.2f us %.3f" % (r["workload"], r["kernel"], r["us_per_launch"], r["roofline_frac"]))
PY
echo "== reference arm x$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 2>> $OUT/bench_${N}.err | tee $OUT/bench_ref_${N}.json | cut -c1-300
