"""RockSample's per-rock belief side-statistics (rock.py:78-86, 177-191) in batched mode:
the kernel ``pomdp_rock_belief_update`` against sequences recorded from the UNMODIFIED reference
(tests/golden/rock_stats.npz, oracle/gen_rollouts.py) -- float64 products compared bit for bit, including the
underflow to 0 and the NaN that the reference's own 0/0 produces -- and against the oracle's restatement."""
import numpy as np
import pytest
import torch

import gym_pomdp_b200 as gp
from oracle import pomdp_oracle as O

from backends import backend  # noqa: F401

SEED = 0x5EED
KEYS = ("count", "measured", "lkv", "lkw", "prob_valuable")


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("tag", ["rock_7_8", "rock_15_15", "srock_11_11"])
def test_belief_stats_follow_the_reference(golden, backend, tag):
    g = golden("rock_stats")
    n, k, stoch, T = (int(v) for v in g[tag + "_cfg"])
    M = len(g[tag + "_x0"])
    env = gp.make("StochasticRock-v0" if stoch else "Rock-v0", board_size=n, num_rocks=k, batch_size=M, device=backend, seed=SEED,
                  track_belief_stats=True)
    env._step_ctr = int(g["reset_ctr"]) - 1
    env.reset()
    x0, y0, st0, _ = (v.cpu().numpy() for v in env.unpack(env.state))
    assert same(x0, g[tag + "_x0"]) and same(y0, g[tag + "_y0"]) and same(st0, g[tag + "_st0"])
    alive = np.ones(M, bool)
    T_run = T
    saw_nan = False
    for t in range(T_run):
        alive &= g[tag + "_alive"][:, t]
        if not alive.any():
            break
        a = torch.as_tensor(g[tag + "_acts"][:, t], device=backend).int()
        ob, rw, done, info = env.step(a)
        assert same(ob.cpu().numpy()[alive], g[tag + "_obs"][alive, t])
        st = info["belief_stats"]
        for key in KEYS:
            got = getattr(st, key).cpu().numpy()[alive]
            exp = g[f"{tag}_{key}"][alive, t]
            assert same(got.astype(np.float64), exp), (tag, t, key)
        saw_nan = saw_nan or bool(np.isnan(st.prob_valuable.cpu().numpy()[alive]).any())
        alive &= ~done.cpu().numpy()
    if tag == "rock_7_8":
        assert saw_nan                                   # lkv = lkw = 0 -> the reference's 0/0
    # masked reset restores the fresh values for the selected envs only
    mask = torch.zeros(M, dtype=torch.bool, device=backend)
    mask[::2] = True
    before = env.belief_stats.lkv.clone()
    env.reset(mask=mask)
    s = env.belief_stats
    assert (s.count[mask] == 0).all() and (s.measured[mask] == 0).all() and (s.lkv[mask] == 1).all() and (s.prob_valuable[mask] == .5).all()
    assert torch.equal(s.lkv[~mask].nan_to_num(-1), before[~mask].nan_to_num(-1))


def test_belief_update_matches_oracle_on_random_batches(backend):
    N, k, board = 4000, 11, 11
    env = gp.make("Rock-v0", board_size=board, num_rocks=k, batch_size=N, device=backend, seed=3)
    cfg = O.RockCfg(board, k)
    rs = np.random.RandomState(0)
    x, y, status = rs.randint(0, board, N), rs.randint(0, board, N), rs.randint(-1, 2, (N, k))
    state = env.pack(x, y, status)
    stats = env.new_belief_stats(N)
    side = [[dict(count=0, measured=0, lkv=1., lkw=1., prob_valuable=.5) for _ in range(k)] for _ in range(N)]
    for t in range(6):
        action = rs.randint(0, 16, N)
        ob = rs.randint(0, 3, N)                           # includes obs 0 (no reading) and non-check actions: untouched
        env.update_belief_stats(stats, state, torch.as_tensor(action, device=backend).int(), torch.as_tensor(ob, device=backend).int())
        for i in range(0, N, 5):
            O.rock_belief_update(cfg, int(x[i]), int(y[i]), int(action[i]), int(ob[i]), side[i])
    for key in KEYS:
        got = getattr(stats, key).cpu().numpy()
        for i in range(0, N, 5):
            assert same(got[i].astype(np.float64), np.array([r[key] for r in side[i]], np.float64)), (key, i)
