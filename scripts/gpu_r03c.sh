#!/bin/bash
# r03c: the new bench line (configs table, collective row, real-reference cpu leg, median-of-replays), the reference arm,
# the ncu launch list of the same command, and the instruction counts behind Network's integer-issue roofline
TAG=${1:-r03c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== bench"; timeout 900 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-600
tail -5 $OUT/bench.err
echo "== bench --steps 20 (what the driver runs)"; timeout 900 python bench.py --steps 20 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_k20.json | cut -c1-400
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_reference.json | cut -c1-600
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --no-graph --no-cpu --no-configs --profiler-range --steps 20 --warmup 3 --e2e-steps 1 > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu issue counts (Network step, Tag step)"
timeout 600 ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second \
    --clock-control none -k regex:pomdp_step_kernel -c 6 --csv --log-file $OUT/issue_network.csv \
    python scripts/bench_configs.py --quick --no-rollout --only "Network-v0" > $OUT/ncu_issue_network.log 2>&1
timeout 600 ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second \
    --clock-control none -k regex:pomdp_step_kernel -c 6 --csv --log-file $OUT/issue_tag.csv \
    python scripts/bench_configs.py --quick --no-rollout --only "Tag-v0 B=2^22" > $OUT/ncu_issue_tag.log 2>&1
ls -la $OUT
