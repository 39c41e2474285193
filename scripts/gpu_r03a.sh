#!/bin/bash
# r03a: GPU parity tests + every config's kernels after the reset rework (Rock one-word reset, Tag digit reset, BattleShip tables)
TAG=${1:-r03a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== configs"; timeout 900 python scripts/bench_configs.py --no-rollout --out $OUT/configs.json 2> $OUT/configs.err | tail -60
tail -5 $OUT/configs.err
