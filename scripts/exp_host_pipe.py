#!/usr/bin/env python
"""Chunk size / slot count sweep of the C host pipeline (pomdp_step_packed_host) for RockSample(11,11), 2^22 envs.
python scripts/exp_host_pipe.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402

dev = torch.device("cuda", 0)
B = 1 << 22
env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=B, device=dev, seed=1)
g = torch.Generator(device=dev); g.manual_seed(0)
s = env.pack(torch.randint(0, 11, (B,), generator=g, device=dev), torch.randint(0, 11, (B,), generator=g, device=dev),
             torch.randint(-1, 2, (B, 11), generator=g, device=dev))
a = torch.randint(0, 16, (B,), generator=g, device=dev, dtype=torch.int32)
hs, ha = s.cpu().pin_memory(), a.cpu().pin_memory()
pin = dict(device="cpu", pin_memory=True)
hp = (torch.empty(B, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.int32, **pin))
res = {}
for ns in (2, 3, 4, 6):
    for lg in (15, 16, 17, 18, 19, 20):
        kw = dict(step_ctr=1, packed=True, pipeline="c", chunk=1 << lg, n_streams=ns)
        for _ in range(3):
            env.simulate_host(hs, ha, hp, **kw)
        t0 = time.perf_counter()
        for _ in range(20):
            env.simulate_host(hs, ha, hp, **kw)
        dt = (time.perf_counter() - t0) / 20
        res["slots=%d chunk=2^%d" % (ns, lg)] = round(dt * 1e3, 3)
t0 = time.perf_counter()
for _ in range(20):
    env.simulate_host(hs, ha, hp, step_ctr=1, packed=True, pipeline="python")
res["python pipeline (2^20, 3 streams)"] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
print(json.dumps(res, indent=0))
best = min(res, key=res.get)
print("best:", best, res[best], "ms ->", B / res[best] / 1e-3, "env-steps/s")
