#!/usr/bin/env python
"""Latency of the scalar (drop-in, batch_size=None) mode: microseconds per env.step()."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_pomdp_b200 as gp  # noqa: E402

out = {}
for env_id, kw in [("Rock-v0", dict(board_size=11, num_rocks=11)), ("Tag-v0", {}), ("Tiger-v0", {}), ("Network-v0", {}),
                   ("Battleship-v0", dict(board_size=(10, 10)))]:
    env = gp.make(env_id, device="cuda:0", seed=1, **kw)
    env.reset()
    n, t0 = 0, time.perf_counter()
    acts = {"Rock-v0": [5, 6, 7, 8], "Tag-v0": [0, 1, 2, 3], "Tiger-v0": [2], "Network-v0": [20, 0, 3], "Battleship-v0": list(range(60))}[env_id]
    for i in range(300):
        ob, rw, done, info = env.step(acts[i % len(acts)])
        n += 1
        if done:
            env.reset()
    out[env_id] = round((time.perf_counter() - t0) / n * 1e6, 1)
print(json.dumps({"scalar_step_us": out}))
