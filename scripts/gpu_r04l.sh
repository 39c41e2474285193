#!/bin/bash
# r04l: Network step specialised on two machine groups (no uniform branches, interleaved Philox chains) -- parity + timing
OUT=gpurun_out/r04l; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "network or Network or rollout or fullsize or edge or gym_surface or parity or stochastic" 2>&1 | tail -3 | tee $OUT/pytest.log
python scripts/bench_configs.py --only "Network" --kernels step,step_packed,rollout --out $OUT/configs.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%-30s %-14s %8.2f us  %.3f of peak' % (d['config'], d['kernel'], d['us_per_launch'], d['frac_of_peak']))
" | tee $OUT/configs.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:pomdp_step_kernel -c 6 --csv --log-file $OUT/issue_network.csv \
  python scripts/bench_configs.py --no-rollout --only "Network" --quick > $OUT/ncu.log 2>&1
grep -E "inst_executed|issue_active|time_duration" $OUT/issue_network.csv | awk -F'","' '{print $5, $13, $15}' | sed 's/(Params.*)//' | sort | uniq | head -12
