#!/bin/bash
# r04b: Network step with the joint failure draw -- GPU parity (all tests), timings, issue metrics, one full ncu capture
OUT=gpurun_out/r04b; mkdir -p $OUT
echo "== pytest -m gpu"; POMDP_DIST_REPORT=$OUT/chisq.jsonl timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
python scripts/bench_configs.py --only "Network" --out $OUT/configs.json 2>&1 | tail -8 | cut -c1-260 | tee $OUT/configs.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  --clock-control none -k regex:pomdp_step_kernel -c 6 --csv --log-file $OUT/issue_network.csv \
  python scripts/bench_configs.py --no-rollout --only "Network" --quick > $OUT/ncu.log 2>&1
grep -c pomdp_ $OUT/issue_network.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pomdp_step_kernel -s 2 -c 1 -o $OUT/network_step_full \
  python scripts/bench_configs.py --no-rollout --only "Network" --quick > $OUT/ncu_full.log 2>&1
ls -la $OUT
