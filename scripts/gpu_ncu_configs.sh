#!/bin/bash
# ncu metric capture (CSV only, no .ncu-rep: small) of every kernel family, one launch each.
# Usage: gpurun --timeout 1500 -- bash scripts/gpu_ncu_configs.sh <tag>
TAG=${1:-ncu}; OUT=gpurun_out/$TAG; mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,launch__shared_mem_per_block_static,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
for pair in "RockSample(11,11):rock11" "RockSample(15,15) B=2^22:rock15" "Tag-v0 B=2^20:tag" "BattleShip 10x10 B=2^18:battleship" "Network-v0:network" "Tiger-v0:tiger"; do
  label="${pair%%:*}"; short="${pair##*:}"
  echo "== ncu $short"
  # -s skips the warm-up launches; each kernel family appears several times, keep a handful of every name
  timeout 600 ncu --metrics $M --clock-control none -k regex:pomdp_ --csv --log-file $OUT/ncu_$short.csv \
      python scripts/bench_configs.py --quick --only "$label" > $OUT/ncu_$short.log 2>&1
  grep -c pomdp_ $OUT/ncu_$short.csv
done
ls -la $OUT
