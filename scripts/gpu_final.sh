#!/bin/bash
# final check of the committed state: full GPU suite, smoke, one bench line (both arms are in r04t)
OUT=gpurun_out/final; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 2> $OUT/bench.err > $OUT/bench_k20.json; tail -2 $OUT/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final/bench_k20.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["pattern_roof"]["step_vs_roof"], d["e2e"]["value"], d["cpu_baseline"]["kind"], d["collective"]["hist_plus_allreduce_us"])
PY
