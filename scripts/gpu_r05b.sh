#!/bin/bash
# r05b: Tag's whole-transition LUT (one table word per env-step) on top of the one-word draw -- the whole GPU suite, Tag's
# config rows, the instruction counts behind Tag's issue roofline, compute-sanitizer over the Tag tests.
#   gpurun -- bash scripts/gpu_r05b.sh
OUT=gpurun_out/r05b; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== Tag configs"; timeout 300 python scripts/bench_configs.py --only Tag --kernels step,step_packed,rollout --out $OUT/tag_configs.json 2>&1 | tail -8
echo "== Tag shared agent cell / single action"; timeout 300 python scripts/bench_action_classes.py --only Tag --out $OUT/tag_action_classes.json 2>&1 | tail -8
echo "== ncu issue counts (Tag step)"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
    --clock-control none -k regex:pomdp_step_kernel -c 6 --csv --log-file $OUT/issue_tag.csv \
    python scripts/bench_configs.py --quick --no-rollout --kernels step --only "Tag-v0 B=2^22" > $OUT/ncu_issue_tag.log 2>&1
tail -4 $OUT/issue_tag.csv | cut -c1-400
echo "== compute-sanitizer (Tag tests)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_golden.py tests/test_edge_cases.py -m gpu -q -x -k "tag or Tag" 2>&1 | tail -6 | tee $OUT/sanitizer_tag.log
