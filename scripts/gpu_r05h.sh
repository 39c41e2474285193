#!/bin/bash
# r05h: the step kernels with the histogram epilogue (pomdp_E_step_hist, ABI 13): GPU suite, and config 5's whole step on one
# rank at the shard size of the 8-GPU run (2^22): three launches vs two vs ONE.   gpurun -- bash scripts/gpu_r05h.sh
OUT=gpurun_out/r05h; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== config 5's step on one rank, shard size 2^22"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 \
    scripts/bench_fused_hist.py --log2-global 22 --out $OUT/step_hist_1rank_2p22.json 2> $OUT/step_hist.err | cut -c1-3000
tail -3 $OUT/step_hist.err
echo "== compute-sanitizer (step + histogram tests)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_edge_cases.py -m gpu -q -x -k "histogram_epilogue or once" 2>&1 | tail -5 | tee $OUT/sanitizer_step_hist.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_edge_cases.py -m gpu -q -x -k "histogram_epilogue and (rock15 or tag3 or tiger)" 2>&1 | tail -5 | tee -a $OUT/sanitizer_step_hist.log
