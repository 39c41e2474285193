timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29531 scripts/bench_fused_hist.py --log2-global 20 2>&1 | tail -3 | cut -c1-600
