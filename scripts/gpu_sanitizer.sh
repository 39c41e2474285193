#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the small-size GPU tests: every kernel family incl. the TMA /
# mbarrier paths.   gpurun --timeout 1500 -- bash scripts/gpu_sanitizer.sh <tag>
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
T="tests/test_parity_golden.py tests/test_edge_cases.py tests/test_queries.py tests/test_coord_ops.py tests/test_host_pipeline.py tests/test_rollout.py tests/test_heuristics.py tests/test_rock_belief_stats.py"
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 420 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 --log-file $OUT/$tool.log \
      python -m pytest $T -m gpu -q -x -k "not 2p2 and not whole" > $OUT/${tool}_pytest.log 2>&1
  echo "exit $?"; tail -2 $OUT/${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/$tool.log | tail -2
done
