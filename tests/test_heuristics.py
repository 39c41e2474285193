"""Heuristic action sets: ``RockEnv._generate_preferred`` (rock.py:293-374, use_heuristic=True),
``RockEnv._select_target`` (rock.py:389-399) and ``TagEnv._generate_preferred`` (tag.py:231-243) -- the scalar hooks, the
batched mask / policy kernels and the fused ``rollout(policy="preferred")`` -- against episodes PLAYED BY THE UNMODIFIED
REFERENCE (tests/golden/heuristic_rollouts.npz, oracle/gen_rollouts.py: its own History / Transition classes, its own
``np.random.choice(env._generate_preferred(history))`` loop with the draws scripted from the Philox words).

Two history layouts are recorded for RockSample: ``*_fields`` builds every Transition by field name; ``*_main`` builds it
as the reference's own loop does, positionally -- Transition(ob, action, next_ob, rw, done) against the field order
(observation, action, reward, next_observation, done), rock.py:525-530 and 566 -- which puts the reward where the
heuristic reads ``next_observation``.  Both are the caller's business; the kernels take the field as given."""
import types

import numpy as np
import pytest
import torch

import gym_pomdp_b200 as gp
from gym_pomdp_b200.envs.rock import History, Transition

from backends import backend  # noqa: F401

ROCK_TAGS = ["rock_7_8_fields", "rock_7_8_main", "rock_11_11_fields", "rock_11_11_main", "rock_4_3_fields", "srock_7_8_fields"]


def make_rock(g, tag, backend, **kw):
    n, k, stoch, T, positional = (int(v) for v in g[tag + "_cfg"])
    env = gp.make("StochasticRock-v0" if stoch else "Rock-v0", board_size=n, num_rocks=k, use_heuristic=True, device=backend,
                  seed=int(g["seed"]), **kw)
    return env, n, k, T, bool(positional)


@pytest.mark.parametrize("tag", ROCK_TAGS)
def test_rock_fused_heuristic_rollout_equals_the_reference(golden, backend, tag):
    g = golden("heuristic_rollouts")
    M = len(g[tag + "_ret"])
    env, n, k, T, positional = make_rock(g, tag, backend, batch_size=M)
    state, ob0 = env.init_states(M, step_ctr=int(g["reset_ctr"]))
    x0, y0, st0, _ = (v.cpu().numpy() for v in env.unpack(state))
    assert np.array_equal(x0, g[tag + "_x0"]) and np.array_equal(y0, g[tag + "_y0"]) and np.array_equal(st0, g[tag + "_st0"])
    stats, hist = env.new_belief_stats(M), env.new_history(M)
    final, ret, steps, flags = env.rollout(state, max_steps=T, step_ctr=int(g["first_ctr"]), policy="preferred", stats=stats,
                                           history=hist, next_is_reward=positional)
    assert np.array_equal(steps.cpu().numpy(), g[tag + "_steps"])
    assert np.array_equal(ret.cpu().numpy(), g[tag + "_ret"])                         # float64, accumulated as CPython does
    x1, y1, st1, done = (v.cpu().numpy() for v in env.unpack(final))
    assert np.array_equal(x1, g[tag + "_x1"]) and np.array_equal(y1, g[tag + "_y1"]) and np.array_equal(st1, g[tag + "_st1"])
    assert np.array_equal(done, g[tag + "_done"])
    # planes passed in are left at their end-of-rollout values; without planes the same rollout starts fresh
    final2, ret2, steps2, _ = env.rollout(state, max_steps=T, step_ctr=int(g["first_ctr"]), policy="preferred",
                                          next_is_reward=positional)
    assert torch.equal(ret, ret2) and torch.equal(final, final2) and torch.equal(steps, steps2)
    last = np.array([g[tag + "_obs"][e, s - 1] if s else 0 for e, s in enumerate(g[tag + "_steps"])])
    assert np.array_equal(hist.prev_obs.cpu().numpy(), last)
    assert int(stats.measured.sum()) == int((g[tag + "_acts"] >= 5).sum() - ((g[tag + "_acts"] >= 5) & (g[tag + "_obs"] == 0)).sum())


@pytest.mark.parametrize("tag", ROCK_TAGS)
def test_rock_preferred_sets_and_draws_step_by_step(golden, backend, tag):
    """mask kernel + policy kernel + step + belief/history updates, one launch each per step: the preferred set, the drawn
    action and the observation of every step of every episode equal the reference's."""
    g = golden("heuristic_rollouts")
    M = len(g[tag + "_ret"])
    env, n, k, T, positional = make_rock(g, tag, backend, batch_size=M, track_history=True, history_next_is_reward=bool(
        int(g[tag + "_cfg"][4])))
    env._step_ctr = int(g["reset_ctr"]) - 1
    env.reset()
    alive = np.ones(M, bool)
    n_fallback = 0
    for t in range(T):
        alive &= g[tag + "_acts"][:, t] >= 0
        if not alive.any():
            break
        words = env.preferred_mask_words().cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        fb = g[tag + "_fallback"][:, t] == 1
        assert np.array_equal(words[alive & ~fb], g[tag + "_mask"][alive & ~fb, t]), (tag, t)
        assert (words[alive & fb] == 0).all(), (tag, t)
        n_fallback += int((alive & fb).sum())
        pref = env._generate_preferred(None).cpu().numpy()                         # bool[M, n_actions]; fallback rows = legal set
        got = (pref * (1 << np.arange(pref.shape[1]))).sum(1)
        assert np.array_equal(got[alive], g[tag + "_mask"][alive, t]), (tag, t)
        env._step_ctr = int(g["first_ctr"]) + t - 1
        a = env.sample_preferred_actions()
        assert np.array_equal(a.cpu().numpy()[alive], g[tag + "_acts"][alive, t]), (tag, t)
        a = torch.where(torch.as_tensor(alive, device=a.device), a, torch.zeros_like(a))
        ob, rw, done, info = env.step(a)
        assert np.array_equal(ob.cpu().numpy()[alive], g[tag + "_obs"][alive, t]), (tag, t)
        alive &= ~done.cpu().numpy()
    assert n_fallback == int((g[tag + "_fallback"] == 1).sum())


def test_rock_scalar_generate_preferred_and_select_target(golden, backend):
    """The scalar hook with the caller's own History object: the list at every step of replayed reference episodes, in both
    layouts; and _select_target on recorded cases."""
    g = golden("heuristic_rollouts")
    for tag in ("rock_7_8_fields", "rock_7_8_main"):
        env, n, k, T, positional = make_rock(g, tag, backend)
        for e in range(6):
            env._step_ctr = int(g["reset_ctr"]) - 1
            env.global_offset = e
            ob = env.reset()
            history = History()
            for t in range(int(g[tag + "_steps"][e])):
                pref = env._generate_preferred(history)
                if g[tag + "_fallback"][e, t]:
                    assert pref == env._generate_legal()
                assert sum(1 << a for a in set(pref)) == int(g[tag + "_mask"][e, t]), (tag, e, t, pref)
                a = int(g[tag + "_acts"][e, t])
                env._step_ctr = int(g["first_ctr"]) + t - 1
                nob, rw, done, info = env.step(a)
                assert nob == int(g[tag + "_obs"][e, t])
                history.append(Transition(ob, a, nob, rw, done) if positional else
                               Transition(observation=ob, action=a, reward=rw, next_observation=nob, done=done))
                ob = nob
        env.close()
    env = gp.make("Rock-v0", board_size=11, num_rocks=11, device=backend)
    for i in range(len(g["select_target"])):
        st = types.SimpleNamespace(agent_pos=(int(g["select_ax"][i]), int(g["select_ay"][i])), rocks=[
            types.SimpleNamespace(status=int(g["select_status"][i, j]), count=int(g["select_count"][i, j]), pos=env._rock_pos[j])
            for j in range(11)])
        assert env._select_target(st, 11) == int(g["select_target"][i]), i
    # without use_heuristic the hook is _generate_legal (rock.py:293-295)
    env.reset()
    assert env._generate_preferred(History()) == env._generate_legal()


@pytest.mark.parametrize("tag", ["tag_1opp", "tag_2opp"])
def test_tag_heuristic_rollouts_equal_the_reference(golden, backend, tag):
    g = golden("heuristic_rollouts")
    n_opp, T = (int(v) for v in g[tag + "_cfg"])
    M = len(g[tag + "_ret"])
    env = gp.make("Tag-v0", num_opponents=n_opp, batch_size=M, device=backend, seed=int(g["seed"]))
    state, ob0 = env.init_states(M, step_ctr=int(g["reset_ctr"]))
    agent, opp, _, _ = (v.cpu().numpy() for v in env.unpack(state))
    assert np.array_equal(agent, g[tag + "_agent0"]) and np.array_equal(opp, g[tag + "_opp0"].reshape(M, n_opp))
    assert np.array_equal(ob0.cpu().numpy(), g[tag + "_ob0"])
    # fused
    lo = torch.zeros(M, dtype=torch.int32, device=backend)
    la = torch.full((M,), -1, dtype=torch.int32, device=backend)
    final, ret, steps, flags = env.rollout(state, max_steps=T, step_ctr=int(g["first_ctr"]), policy="preferred", last_obs=lo,
                                           last_action=la)
    assert np.array_equal(steps.cpu().numpy(), g[tag + "_steps"]) and np.array_equal(ret.cpu().numpy(), g[tag + "_ret"])
    a1, o1, n1, done = (v.cpu().numpy() for v in env.unpack(final))
    assert np.array_equal(a1, g[tag + "_agent1"]) and np.array_equal(n1, g[tag + "_nopp1"]) and np.array_equal(done, g[tag + "_done"])
    assert np.array_equal(o1, g[tag + "_opp1"].reshape(M, n_opp))
    last_t = g[tag + "_steps"] - 1
    assert np.array_equal(la.cpu().numpy(), g[tag + "_acts"][np.arange(M), last_t])
    assert np.array_equal(lo.cpu().numpy(), g[tag + "_obs"][np.arange(M), last_t])
    final2, ret2, _, _ = env.rollout(state, max_steps=T, step_ctr=int(g["first_ctr"]), policy="preferred")   # empty history
    assert torch.equal(ret, ret2) and torch.equal(final, final2)
    # step by step: mask kernel, policy kernel, step
    s = state.clone()
    lo, la = None, None
    alive = np.ones(M, bool)
    for t in range(T):
        alive &= g[tag + "_acts"][:, t] >= 0
        if not alive.any():
            break
        words = env.preferred_mask_words(s, lo, la).cpu().numpy()
        assert np.array_equal(words[alive], g[tag + "_mask"][alive, t]), (tag, t)
        a = env.sample_preferred_actions(s, lo, la, step_ctr=int(g["first_ctr"]) + t)
        assert np.array_equal(a.cpu().numpy()[alive], g[tag + "_acts"][alive, t]), (tag, t)
        a = torch.where(torch.as_tensor(alive, device=a.device), a, torch.zeros_like(a))
        s, ob, rw, fl = env.simulate(s, a, step_ctr=int(g["first_ctr"]) + t)
        assert np.array_equal(ob.cpu().numpy()[alive], g[tag + "_obs"][alive, t]), (tag, t)
        lo, la = ob, a
        alive &= (fl.cpu().numpy() & 1) == 0
    assert (g[tag + "_mask"] == 16).sum() > 0                                          # the [TAG]-only branch was played


def test_tag_scalar_generate_preferred(golden, backend):
    g = golden("heuristic_rollouts")
    tag = "tag_1opp"

    class Hist(list):                                                                  # the History the reference's Tag loop expects
        size = property(len)

        def add(self, action, ob):
            self.append(types.SimpleNamespace(action=action, ob=ob))
    env = gp.make("Tag-v0", device=backend, seed=int(g["seed"]))
    for e in range(8):
        env._step_ctr = int(g["reset_ctr"]) - 1
        env.global_offset = e
        assert env.reset() == int(g[tag + "_ob0"][e])
        h = Hist()
        for t in range(int(g[tag + "_steps"][e])):
            pref = env._generate_preferred(h)
            assert sum(1 << a for a in pref) == int(g[tag + "_mask"][e, t]) and pref == sorted(pref)
            a = int(g[tag + "_acts"][e, t])
            env._step_ctr = int(g["first_ctr"]) + t - 1
            ob, rw, done, info = env.step(a)
            assert ob == int(g[tag + "_obs"][e, t])
            h.add(a, ob)
