#!/usr/bin/env python
"""A bootstrap particle filter + Monte-Carlo action selection on RockSample(11,11), the caller pattern this library
serves: what a planner does with the reference as ``for p in particles: env._set_state(p);
env.step(a)`` is here three kernel launches per real step, whatever the number of particles.

    python examples/rock_particle_filter.py [--particles 1048576] [--steps 40] [--device cuda:0]

Per real step:  for each candidate action, ONE fused rollout launch evaluates it from every particle (first step with
that action, then a uniform-legal rollout: ``rollout(first_action=a)``);  the real env steps;  ``simulate`` moves all particles with the chosen
action, ``observation_prob`` (the reference's ``_compute_prob``) weights them by the real observation, and
``torch.multinomial`` resamples.  ``belief_histogram`` summarises the belief (per-rock "still good" counts).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402


def run(n_particles=1 << 20, steps=40, device="cuda:0", seed=7, rollout_depth=20, verbose=True):
    board, k = 11, 11
    world = gp.make("Rock-v0", board_size=board, num_rocks=k, device=device, seed=seed)                 # the real env: one instance
    sim = gp.make("Rock-v0", board_size=board, num_rocks=k, batch_size=n_particles, device=device, seed=seed + 1)
    world.reset()
    particles, _ = sim.init_states(n_particles)                # belief = the reset distribution (rock statuses ~ Bernoulli(1/2))
    n_actions = sim.action_space.n
    total, discount = 0.0, 1.0
    for t in range(steps):
        # ---- choose: mean discounted return of (action, then uniform-legal rollout) over a subsample of the belief
        sub = particles[torch.randint(0, n_particles, (min(n_particles, 1 << 14),), device=particles.device)]
        legal = sim.legal_mask(sub).float().mean(0) > 0.5
        best, best_q = 1, -1e30
        for a in torch.nonzero(legal)[:, 0].tolist():
            _, ret, _, _ = sim.rollout(sub, max_steps=rollout_depth, first_action=a)     # Q(s, a) samples: one launch
            q = float(ret.mean())
            if q > best_q:
                best, best_q = a, q
        # ---- act in the real env
        ob, rw, done, _ = world.step(best)
        total += discount * rw
        discount *= world._discount
        # ---- belief update: propagate, weight by the real observation, resample
        act = torch.full((n_particles,), best, dtype=torch.int32, device=particles.device)
        nxt, _, _, flags = sim.simulate(particles, act)
        w = sim.observation_prob(act, nxt, torch.full((n_particles,), ob, dtype=torch.int32, device=nxt.device))
        w = w * ((flags & ~1) == 0)                            # drop particles the step flagged as impossible
        if float(w.sum()) <= 0:                                # belief collapsed (cannot happen with exact weights)
            w = torch.ones_like(w)
        particles = nxt[torch.multinomial(w, n_particles, replacement=True)]
        if verbose:
            hist = sim.belief_histogram(particles)
            print("t=%2d a=%2d ob=%d rw=%4d  P(rock good)=%s" % (t, best, ob, rw, " ".join("%.2f" % (float(h) / n_particles)
                                                                                   for h in hist[:k])))
        if done:
            break
    return total, t + 1


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args()
    ret, n = run(a.particles, a.steps, a.device)
    print("discounted return %.2f over %d steps" % (ret, n))
