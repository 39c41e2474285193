#!/bin/bash
# r04j: heuristic rollouts with the rule's per-rock predicates as register bit masks -- parity + timings
OUT=gpurun_out/r04j; mkdir -p $OUT
timeout 900 python -m pytest tests/test_heuristics.py tests/test_rollout.py -m gpu -q -x 2>&1 | tail -3 | tee $OUT/pytest.log
python scripts/bench_heuristic_rollouts.py --out $OUT/heuristic_rollouts.json | tee $OUT/heur.log
