#!/bin/bash
# r05l: the local histogram ending as ONE atomic per bin per CTA (count + arrival in one word).   gpurun -- bash scripts/gpu_r05l.sh
OUT=gpurun_out/r05l; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
echo "== histograms of every config"; timeout 600 python scripts/bench_configs.py --kernels belief_hist --out $OUT/hist_configs.json 2>&1 | tail -12 | cut -c1-200
echo "== config 5's step on one rank, shard size 2^22"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 \
    scripts/bench_fused_hist.py --log2-global 22 --out $OUT/step_hist_1rank_2p22.json 2> $OUT/step_hist.err | cut -c1-1600
echo "== compute-sanitizer racecheck + memcheck (histogram tests)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_edge_cases.py -m gpu -q -x -k "histogram_epilogue or once or bincount" 2>&1 | tail -3 | tee $OUT/sanitizer_hist.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_edge_cases.py -m gpu -q -x -k "once or bincount" 2>&1 | tail -3 | tee -a $OUT/sanitizer_hist.log
