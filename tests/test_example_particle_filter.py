"""The end-to-end caller pattern (examples/rock_particle_filter.py): rollouts for action selection, simulate +
observation_prob + resampling for the belief update, belief_histogram for the summary -- every batched entry point a
POMCP / particle-filter caller touches, exercised together."""
import os
import sys

import torch

from backends import backend  # noqa: F401

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


def test_particle_filter_runs_and_tracks_the_truth(backend):
    import rock_particle_filter as pf
    torch.manual_seed(0)
    n = (1 << 16) if backend.startswith("cuda") else 2048
    ret, steps = pf.run(n_particles=n, steps=12, device=backend, seed=11, rollout_depth=8, verbose=False)
    assert steps >= 1 and -200.0 <= ret <= 200.0
