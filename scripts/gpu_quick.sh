#!/bin/bash
# Quick GPU check: a subset of tests + selected config timings.  Usage: gpurun -- bash scripts/gpu_quick.sh <tag> "<pytest args>" "<only-filter>"
TAG=${1:-quick}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest $2"; timeout 900 python -m pytest $2 -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest.log
if [ -n "$3" ]; then
  echo "== configs ($3)"; timeout 600 python scripts/bench_configs.py --only "$3" --out $OUT/configs.json 2>&1 | tail -12
fi
