#!/bin/bash
# r04a: Network step with the joint failure draw (alias table) -- timings + instruction counts (no parity yet)
OUT=gpurun_out/r04a; mkdir -p $OUT
python scripts/bench_configs.py --no-rollout --only "Network" --out $OUT/configs.json 2>&1 | tail -8 | tee $OUT/configs.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  --clock-control none -k regex:pomdp_step_kernel -c 6 --csv --log-file $OUT/issue_network.csv \
  python scripts/bench_configs.py --no-rollout --only "Network" --quick > $OUT/ncu.log 2>&1
tail -8 $OUT/issue_network.csv
