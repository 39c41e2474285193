#!/usr/bin/env python
"""Config 5's collective on N GPUs (torchrun): belief histogram + all-reduce, three ways, eager and as a CUDA graph of 20 calls
(the graph removes the interpreter from between the launches, so the GPU-side cost of each variant shows):
  hist        the histogram kernel alone (local counts)
  nccl        histogram kernel (self-cleaning) + ncclAllReduce         belief_histogram(all_reduce=True)
  fused       ONE kernel: histogram + reductions into every rank's buffer over NVLink peer memory + arrive/wait,
              then the copy that hands the counts out               belief_histogram(all_reduce="fused")
and the whole step of config 5 -- transition, counts of the next states, sum over the ranks -- as three launches
(step, histogram, ncclAllReduce), two (step, fused histogram) and ONE (simulate_hist(all_reduce="fused"): pomdp_rock_step_hist).
    torchrun --nproc-per-node N scripts/bench_fused_hist.py [--out file.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--log2-global", type=int, default=25)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dist.barrier()
G = 1 << args.log2_global
B = G // world
env = gp.make("Rock-v0", board_size=15, num_rocks=15, batch_size=B, device=dev, seed=0x5EED, global_offset=rank * B)
state, _ = env.init_states(B, step_ctr=1)
ref = env.belief_histogram(state, all_reduce=True)
fused = env.belief_histogram(state, all_reduce="fused")
ok = bool(torch.equal(ref, fused)) and int(ref[15:].sum()) == G
variants = {"hist": lambda: env.belief_histogram(state), "nccl": lambda: env.belief_histogram(state, all_reduce=True),
            "fused": lambda: env.belief_histogram(state, all_reduce="fused")}
# the whole config-5 step: transition of the shard, counts of the next states, sum over the ranks
action = torch.randint(0, 20, (B,), device=dev, dtype=torch.int32)
sout = (torch.empty_like(state), torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.float32, device=dev),
        torch.empty(B, dtype=torch.int32, device=dev))


def step_then(all_reduce):
    env.simulate(state, action, out=sout, step_ctr=7)
    return env.belief_histogram(sout[0], all_reduce=all_reduce)


step_ref = step_then(True)
ok = ok and bool(torch.equal(step_then("fused"), step_ref)) and int(step_ref[15:].sum()) == G
ok = ok and bool(torch.equal(env.simulate_hist(state, action, out=sout, step_ctr=7, all_reduce="fused")[4], step_ref))
ok = ok and bool(torch.equal(env.simulate_hist(state, action, out=sout, step_ctr=7, all_reduce=True)[4], step_ref))
step_variants = {"step": lambda: env.simulate(state, action, out=sout, step_ctr=7)[0],
                 "step_then_hist_nccl": lambda: step_then(True),                   # three launches: step, histogram, ncclAllReduce
                 "step_then_hist_fused": lambda: step_then("fused"),               # two: step, histogram + all-reduce
                 "step_hist_local": lambda: env.simulate_hist(state, action, out=sout, step_ctr=7)[4],      # one, no collective
                 "step_hist_fused": lambda: env.simulate_hist(state, action, out=sout, step_ctr=7, all_reduce="fused")[4]}   # ONE
REPS = 20
res = {}
for name, fn in list(variants.items()) + list(step_variants.items()):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    row = {"eager_us": e0.elapsed_time(e1) * 1e3 / REPS}
    try:
        stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(stream):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for _ in range(REPS):
                    out = fn()
            g.replay()
            torch.cuda.synchronize()
            dist.barrier()
            times = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                g.replay()
                b.record(stream)
                torch.cuda.synchronize()
                times.append(a.elapsed_time(b) * 1e3 / REPS)
        row["graph_us"] = sorted(times)[len(times) // 2]
        if name in ("nccl", "fused"):
            ok = ok and bool(torch.equal(out, ref))
        elif name in ("step_then_hist_nccl", "step_then_hist_fused", "step_hist_fused"):
            ok = ok and bool(torch.equal(out, step_ref))
    except Exception as e:  # noqa: BLE001
        row["graph_error"] = repr(e)[:160]
    t = torch.tensor([row["eager_us"], row.get("graph_us", -1.0)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    row["eager_us"], gmax = float(t[0]), float(t[1])
    if "graph_us" in row:
        row["graph_us"] = gmax
    res[name] = row
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    out = {"workload": "RockSample(15,15) belief histogram, global batch 2^%d over %d rank(s)" % (args.log2_global, world),
           "batch_per_gpu": B, "bins": int(ref.numel()), "all_variants_equal": bool(flag.item() == 1.0),
           "us_per_call_max_over_ranks": {k: v for k, v in res.items() if k in variants},
           "step_pipeline_us_max_over_ranks": {k: v for k, v in res.items() if k in step_variants}}
    print(json.dumps(out))
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(json.dumps(out, indent=1))
dist.destroy_process_group()
