#!/bin/bash
# r05a: Tag's one-word opponent draw (ABI 11) -- the whole GPU suite, Tag's config rows, and the belief histogram on batches
# whose particles share the agent's cell.   gpurun -- bash scripts/gpu_r05a.sh
OUT=gpurun_out/r05a; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== Tag configs"; timeout 300 python scripts/bench_configs.py --only Tag --kernels step,step_packed,rollout --out $OUT/tag_configs.json 2>&1 | tail -8
echo "== Rock(7,8) + probe"; timeout 300 python scripts/bench_configs.py --only "RockSample(7,8)" --kernels step --out $OUT/rock78_configs.json 2>&1 | tail -3
echo "== histogram, shared agent cell"; timeout 300 python scripts/bench_hist_shared_cell.py --out $OUT/hist_shared_cell.json 2>&1 | tail -14
