#!/usr/bin/env python
"""BASELINE config 5 under torchrun: RockSample(15,15), global batch 2^25 index-sharded over the ranks, one step,
then the belief histogram summed over ranks with NCCL (the only collective on the path).  Rank 0 checks the
all-reduced histogram against the sum of the per-rank ones and prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    B = (1 << 25) // world
    env = gp.make("Rock-v0", board_size=15, num_rocks=15, batch_size=B, device=dev, seed=0x5EED, global_offset=rank * B)
    env.reset()
    g = torch.Generator(device=dev)
    g.manual_seed(rank)
    action = torch.randint(0, 20, (B,), generator=g, device=dev, dtype=torch.int32)
    env.step(action)
    local_hist = env.belief_histogram()
    for _ in range(3):
        env.belief_histogram(all_reduce=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        total = env.belief_histogram(all_reduce=True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros_like(local_hist) for _ in range(world)]
    dist.all_gather(gathered, local_hist)
    ok = bool(torch.equal(total, torch.stack(gathered).sum(0))) and int(total[15:].sum()) == B * world
    if rank == 0:
        print(json.dumps({"check": "belief histogram all-reduce (NCCL)", "env": "RockSample(15,15)", "global_batch": B * world,
                          "n_gpus": world, "bins": int(total.numel()), "hist_plus_allreduce_ms": float(ms.item()), "ok": ok,
                          "rocks_still_good": [int(v) for v in total[:15].tolist()]}))
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
