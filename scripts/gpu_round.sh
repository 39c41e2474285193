#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + full capture of the step kernel.
# Usage (from the CPU box):  gpurun --timeout 1500 -- bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json
tail -5 $OUT/bench.err
echo "== bench eager"; timeout 600 python bench.py --no-graph --no-cpu --steps 500 2>> $OUT/bench.err | tee $OUT/bench_eager.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --no-graph --no-cpu --profiler-range --steps 20 --warmup 3 --e2e-steps 1 > $OUT/ncu_launch_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pomdp_step_kernel -s 5 -c 3 -f -o $OUT/rock_step \
    python bench.py --no-graph --no-cpu --profiler-range --steps 20 --warmup 3 --e2e-steps 1 > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
