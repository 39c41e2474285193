#!/bin/bash
# r04n: bench with the live pattern roof (pomdp_stream_probe) -- default K and the driver's --steps 20
OUT=gpurun_out/r04n; mkdir -p $OUT
timeout 300 python -m pytest tests/test_abi.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --no-cpu --no-configs 2> $OUT/bench.err > $OUT/bench.json; tail -2 $OUT/bench.err
timeout 600 python bench.py --no-cpu --no-configs --steps 20 --warmup 3 2>> $OUT/bench.err > $OUT/bench_k20.json
python - <<'PY'
import json
for f in ("bench.json", "bench_k20.json"):
    d = json.loads(open("gpurun_out/r04n/" + f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["pattern_roof"], d["timing"]["mode"][:40])
PY
