#!/bin/bash
# r05m: the round's evidence run -- GPU parity suite, smoke, bench (both arms, eager), ncu launch list + full capture of the
# headline kernel, every config's timings, per-kernel ncu metrics.
TAG=r05m; OUT=gpurun_out/$TAG; mkdir -p $OUT
bash scripts/gpu_round.sh $TAG
echo "== configs"; timeout 900 python scripts/bench_configs.py --out $OUT/configs.json 2> $OUT/configs.err | tail -3 | cut -c1-200
bash scripts/gpu_ncu_configs.sh $TAG > $OUT/ncu_configs.log 2>&1
tail -3 $OUT/ncu_configs.log
echo "== action classes"; timeout 600 python scripts/bench_action_classes.py --out $OUT/action_classes.json 2>&1 | tail -2 | cut -c1-200
echo "== heuristic rollouts"; timeout 300 python scripts/bench_heuristic_rollouts.py --out $OUT/heuristic_rollouts.json 2>&1 | tail -3 | cut -c1-200
