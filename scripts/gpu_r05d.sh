#!/bin/bash
# r05d: Tag step4 in two flavours (branch-out for the step kernels, fix-up for the rollouts).   gpurun -- bash scripts/gpu_r05d.sh
OUT=gpurun_out/r05d; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== Tag configs"; timeout 300 python scripts/bench_configs.py --only Tag --kernels step,step_packed,rollout --out $OUT/tag_configs.json 2>&1 | tail -8 | cut -c1-330
