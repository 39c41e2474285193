#!/usr/bin/env python
"""Per-action-class timings of the step kernels ("all-move / all-check ... since divergence differs").

    python scripts/bench_action_classes.py [--out gpurun_out/<tag>/action_classes.json]

Same timing as scripts/bench_configs.py (K launches in one CUDA graph over rotating buffer sets > L2), but every launch's
action array holds ONE class of action -- so a warp never diverges between classes -- next to the uniform mix the bench uses.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import torch  # noqa: E402

import bench_configs as BC  # noqa: E402
import gym_pomdp_b200 as gp  # noqa: E402

CASES = [
    # name, id, kwargs, log2 B, label, {class: (lo, hi) of the action range}
    ("rock", "Rock-v0", dict(board_size=11, num_rocks=11), 22, "RockSample(11,11) B=2^22",
     {"uniform": (0, 16), "move (0-3)": (0, 4), "east only (1)": (1, 2), "sample (4)": (4, 5), "check (5-15)": (5, 16),
      "check rock 3 only (8)": (8, 9)}),
    ("rock", "Rock-v0", dict(board_size=7, num_rocks=8), 20, "RockSample(7,8) B=2^20",
     {"uniform": (0, 13), "sample (4)": (4, 5), "check rock 3 only (8)": (8, 9)}),
    ("rock", "Rock-v0", dict(board_size=15, num_rocks=15), 22, "RockSample(15,15) B=2^22",
     {"uniform": (0, 20), "sample (4)": (4, 5), "check rock 3 only (8)": (8, 9)}),
    ("rock", "StochasticRock-v0", dict(board_size=11, num_rocks=11), 22, "StochasticRock(11,11) B=2^22",
     {"uniform": (0, 16), "move (0-3)": (0, 4), "check (5-15)": (5, 16)}),
    ("tag", "Tag-v0", {}, 22, "Tag-v0 B=2^22", {"uniform": (0, 5), "move (0-3)": (0, 4), "tag (4)": (4, 5)}),
    ("tiger", "Tiger-v0", {}, 22, "Tiger-v0 B=2^22", {"uniform": (0, 3), "open (0-1)": (0, 2), "listen (2)": (2, 3)}),
    ("network", "Network-v0", {}, 22, "Network-v0 B=2^22",
     {"uniform": (0, 21), "ping / reboot (0-19)": (0, 20), "no-op (20)": (20, 21)}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="substring filter on the label")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak, peak_src = BC.peak_gbs()
    rows = []
    for name, env_id, kw, lg, label, classes in CASES:
        if args.only and args.only not in label:
            continue
        B = 1 << lg
        env = gp.make(env_id, batch_size=B, device=dev, seed=0x5EED, **kw)
        nbytes = 8 * env.state_words + 16
        n_sets = max(2, -(-3 * BC.L2 // (B * nbytes)))
        gen = torch.Generator(device=dev)
        gen.manual_seed(0x5EED)
        states = [BC.synth(env, name, B, gen, dev)[0] for _ in range(n_sets)]
        # the agent's own position is observed in RockSample and Tag: the particles of a belief share it
        shared = None
        if name == "rock":
            k = env.num_rocks
            shared = [env.pack(torch.full((B,), 3, device=dev), torch.full((B,), 5, device=dev),
                               torch.randint(-1, 2, (B, k), generator=gen, device=dev)) for _ in range(n_sets)]
        elif name == "tag":
            shared = [env.pack(torch.full((B,), 7, device=dev), torch.randint(0, 29, (B, 1), generator=gen, device=dev))
                      for _ in range(n_sets)]
        outs = [(torch.empty_like(states[0]), torch.empty(B, dtype=torch.int32, device=dev),
                 torch.empty(B, dtype=torch.float32, device=dev), torch.empty(B, dtype=torch.int32, device=dev)) for _ in range(n_sets)]
        for cls, (lo, hi) in classes.items():
            acts = [torch.randint(lo, hi, (B,), generator=gen, device=dev, dtype=torch.int32) for _ in range(n_sets)]

            def step(i):
                env.simulate(states[i % n_sets], acts[i % n_sets], out=outs[i % n_sets], step_ctr=i + 1)
            ms = BC.time_graph(step, args.steps, dev)
            gbs = B * nbytes / (ms * 1e-3) / 1e9
            row = {"config": label, "actions": cls, "us_per_launch": ms * 1e3, "units_per_s": B / (ms * 1e-3), "frac_of_peak": gbs / peak}
            rows.append(row)
            print(json.dumps(row), flush=True)
            if shared is not None and cls != "uniform" and (hi - lo == 1 or name == "tag"):
                def step_shared(i):
                    env.simulate(shared[i % n_sets], acts[i % n_sets], out=outs[i % n_sets], step_ctr=i + 1)
                ms = BC.time_graph(step_shared, args.steps, dev)
                row = {"config": label, "actions": cls + ", one agent cell for the whole batch", "us_per_launch": ms * 1e3,
                       "units_per_s": B / (ms * 1e-3), "frac_of_peak": B * nbytes / (ms * 1e-3) / 1e9 / peak}
                rows.append(row)
                print(json.dumps(row), flush=True)
        del env, states, outs
        torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump({"peak_gbs": peak, "peak_source": peak_src, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
