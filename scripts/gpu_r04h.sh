#!/bin/bash
# r04h: Rock LUT with an odd, skewed row pitch (same-action batches) -- full GPU parity, per-action-class timings, bench
OUT=gpurun_out/r04h; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_gpu.log
python scripts/bench_action_classes.py --out $OUT/action_classes.json 2>&1 | cut -c1-170 | tee $OUT/action_classes.log
timeout 600 python bench.py --no-cpu 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-300
