#!/bin/bash
# r05k: Tiger's one-word draw (ABI 14): GPU suite + Tiger's rows.   gpurun -- bash scripts/gpu_r05k.sh
OUT=gpurun_out/r05k; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
echo "== Tiger configs"; timeout 300 python scripts/bench_configs.py --only Tiger --kernels step,step_packed,rollout --out $OUT/tiger_configs.json 2>&1 | tail -4 | cut -c1-400
