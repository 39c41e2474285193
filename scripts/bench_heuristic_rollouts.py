#!/usr/bin/env python
"""Fused rollouts under _generate_preferred (use_heuristic=True) vs the uniform-legal policy: RockSample(11,11) and Tag-v0,
2^20 envs x 32 steps.   python scripts/bench_heuristic_rollouts.py [--out gpurun_out/<tag>/heuristic_rollouts.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
args = ap.parse_args()
dev = torch.device("cuda", 0)
res = {}
for name, mk in (("Rock(11,11) heuristic", lambda B: gp.make("Rock-v0", board_size=11, num_rocks=11, use_heuristic=True, batch_size=B, device=dev, seed=1)),
                 ("Tag-v0 heuristic", lambda B: gp.make("Tag-v0", batch_size=B, device=dev, seed=1))):
    B = 1 << 20
    env = mk(B)
    s, _ = env.init_states(B, step_ctr=1)
    for pol in ("legal", "preferred"):
        for _ in range(2):
            out = env.rollout(s, max_steps=32, step_ctr=5, policy=pol)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = env.rollout(s, max_steps=32, step_ctr=5, policy=pol)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        n = int(out[2].sum())
        res["%s policy=%s" % (name, pol)] = {"ms_per_launch": ms, "env_steps": n, "env_steps_per_s": n / ms * 1e3}
print(json.dumps(res, indent=1))
if args.out:
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write(json.dumps(res, indent=1))
