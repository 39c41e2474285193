#!/usr/bin/env python
"""Belief histogram on the particles of a REAL belief: the agent's own cell is observed in RockSample and Tag, so every
particle of a belief shares it and the categorical bin they hit is one and the same -- the worst case for per-particle
shared-memory atomics (a 32-way same-address conflict per warp).  Times the histogram over synthetic batches with
uniform agent cells (what bench_configs.py uses) and with one shared agent cell.

    python scripts/bench_hist_shared_cell.py [--out gpurun_out/<tag>/hist_shared_cell.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import torch  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402
from bench_configs import L2, peak_gbs, synth, time_graph  # noqa: E402

CASES = [
    ("rock", "Rock-v0", dict(board_size=7, num_rocks=8), 20, "RockSample(7,8) B=2^20"),
    ("rock", "Rock-v0", dict(board_size=11, num_rocks=11), 22, "RockSample(11,11) B=2^22"),
    ("rock", "Rock-v0", dict(board_size=15, num_rocks=15), 22, "RockSample(15,15) B=2^22"),
    ("rock", "Rock-v0", dict(board_size=15, num_rocks=15), 25, "RockSample(15,15) B=2^25"),
    ("tag", "Tag-v0", {}, 20, "Tag-v0 B=2^20"),
    ("tag", "Tag-v0", {}, 22, "Tag-v0 B=2^22"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak, _ = peak_gbs()
    rows = []
    for name, env_id, kw, lg, label in CASES:
        B = 1 << lg
        env = gp.make(env_id, batch_size=B, device=dev, seed=0x5EED, **kw)
        W = env.state_words
        n_sets = max(2, -(-2 * L2 // (B * 4 * W)))
        gen = torch.Generator(device=dev)
        gen.manual_seed(0x5EED)
        for cells in ("uniform", "shared"):
            sets = []
            for _ in range(n_sets):
                s, _a = synth(env, name, B, gen, dev)
                if cells == "shared":
                    if name == "rock":
                        x, y, st = env.unpack(s)[:3]
                        s = env.pack(torch.full_like(x, 3), torch.full_like(y, 2), st)
                    else:
                        ag, opp = env.unpack(s)[:2]
                        s = env.pack(torch.full_like(ag, 7), opp)
                sets.append(s)
            ref = env.belief_histogram(sets[0]).clone()

            def hist(i):
                env.belief_histogram(sets[i % n_sets])
            K = max(20, min(args.steps, int(args.steps * (1 << 22) / B)))
            ms = time_graph(hist, K, dev)
            gbs = B * 4 * W / (ms * 1e-3) / 1e9
            row = {"config": label, "agent_cells": cells, "us_per_launch": ms * 1e3, "achieved_gbs": gbs, "frac_of_peak": gbs / peak,
                   "total": int(ref.sum().item())}
            rows.append(row)
            print(json.dumps(row), flush=True)
            del sets
        del env
        torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump({"rows": rows, "peak_gbs": peak}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
