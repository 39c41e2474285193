#!/bin/bash
# r05c: one-launch belief histogram (pomdp_belief_hist_once, ABI 12) + Tag's step with the fix-ups under the group test.
#   gpurun -- bash scripts/gpu_r05c.sh
OUT=gpurun_out/r05c; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== Tag configs"; timeout 300 python scripts/bench_configs.py --only Tag --kernels step,step_packed,rollout --out $OUT/tag_configs.json 2>&1 | tail -8 | cut -c1-330
echo "== histograms of every config"; timeout 600 python scripts/bench_configs.py --kernels belief_hist --out $OUT/hist_configs.json 2>&1 | tail -12 | cut -c1-330
echo "== heuristic rollouts"; timeout 300 python scripts/bench_heuristic_rollouts.py --out $OUT/heuristic_rollouts.json 2>&1 | tail -6 | cut -c1-330
