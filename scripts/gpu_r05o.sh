#!/bin/bash
# r05o: the committed state once more -- GPU suite (with the full-size step_hist tests), smoke, the default bench line (both
# arms) with collective.step_pipeline.   gpurun -- bash scripts/gpu_r05o.sh
OUT=gpurun_out/r05o; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py 2> $OUT/bench.err > $OUT/bench.json; tail -2 $OUT/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>> $OUT/bench.err > $OUT/bench_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r05o/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["pattern_roof"]["step_vs_roof"], d["e2e"]["value"], d["cpu_baseline"]["kind"])
print(json.dumps(d["collective"]["step_pipeline"])[:600])
r = json.loads(open("gpurun_out/r05o/bench_reference.json").read().strip().splitlines()[-1]); print("reference arm", r["value"], r["cpu_baseline"]["kind"])
PY
