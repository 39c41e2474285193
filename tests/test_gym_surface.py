"""The drop-in boundary on the Python side: the reference's ``gym.make()``
ids, the old-gym protocol ``reset() -> ob`` / ``step(a) -> (ob, reward, done, {"state": s})``,
``action_space`` / ``observation_space`` / ``_discount``, the reference's asserts, and whole
single-instance EPISODES compared step by step with the oracle's restatement of the
reference under the same Philox words (scalar mode: instance 0, step counter 1, 2, 3 ...).

Runs on the host simulation here and, marked ``gpu``, on the CUDA library.
"""
import numpy as np
import pytest
import torch

import gym_pomdp_b200 as gp
from gym_pomdp_b200.geometry import Coord
from oracle import philox
from oracle import pomdp_oracle as O

from backends import backend  # noqa: F401

SEED = 0xC0FFEE


def words(ctr, domain, n_slots=12):
    w = philox.draw_slots(SEED, np.arange(1), ctr, domain, n_slots)[0]
    return lambda slot: int(w[slot])


def test_registry_ids_match_the_reference():
    # gym_pomdp/__init__.py:7-46 minus Pocman-v0 / Test-v0 (out of scope)
    assert sorted(gp.registry) == sorted(["Tiger-v0", "Tag-v0", "Battleship-v0", "Rock-v0", "StochasticRock-v0",
                                          "Network-v0"])
    with pytest.raises(KeyError):
        gp.make("Pocman-v0")


@pytest.mark.parametrize("env_id,kw,n_act,n_obs,discount", [
    ("Rock-v0", {}, 13, 3, .95),                                       # rock.py:99: board 7, 8 rocks
    ("Rock-v0", dict(board_size=11, num_rocks=11), 16, 3, .95),
    ("Rock-v0", dict(board_size=15, num_rocks=15), 20, 3, .95),
    ("StochasticRock-v0", {}, 13, 3, .95),
    ("Tag-v0", {}, 5, 30, .95),                                        # tag.py:87-95
    ("Battleship-v0", {}, 25, 2, 1.),                                  # battleship.py:67-75: 5x5
    ("Battleship-v0", dict(board_size=(10, 10)), 100, 2, 1.),
    ("Tiger-v0", {}, 3, 3, .95),                                       # tiger.py:49-58
    ("Network-v0", {}, 21, 3, .95),                                    # network.py:27-38
])
def test_spaces_and_protocol(backend, env_id, kw, n_act, n_obs, discount):
    env = gp.make(env_id, device=backend, seed=SEED, **kw)
    assert env.action_space.n == n_act and env.observation_space.n == n_obs
    assert env._discount == discount
    ob = env.reset()
    assert isinstance(ob, int) and env.observation_space.contains(ob)
    a = 0
    out = env.step(a)
    assert isinstance(out, tuple) and len(out) == 4
    ob, rw, done, info = out
    assert isinstance(ob, int) and isinstance(done, bool) and isinstance(rw, (int, float))
    assert set(info) == {"state"}
    # the reference's two asserts at the top of every step()
    with pytest.raises(AssertionError):
        env.step(n_act)
    with pytest.raises(AssertionError):
        env.step(-1)
    assert env.seed(5) == [5]
    env.close()


def test_rock_episode_matches_oracle(backend):
    for board, k in [(7, 8), (11, 11), (15, 15)]:
        env = gp.make("Rock-v0", board_size=board, num_rocks=k, device=backend, seed=SEED)
        cfg = O.RockCfg(board, k)
        assert env.reset() == 0
        x, y, status, _ = O.rock_reset(cfg, words(1, philox.DOMAIN_RESET, 16))
        rs = np.random.RandomState(board)
        ctr = 1
        for _ in range(60):
            a = int(rs.randint(0, 5 + k))
            if board == 15 and (x, y) == (12, 2) and a == 4:
                a = 5
            ctr += 1
            ob, rw, done, info = env.step(a)
            x, y, status, eob, erw, edone, err = O.rock_step(cfg, x, y, status, a, words(ctr, philox.DOMAIN_STEP))
            assert (ob, rw, done) == (eob, erw, edone) and isinstance(rw, int)
            s = info["state"]
            assert s["agent_pos"] == Coord(x, y) and [r["status"] for r in s["rocks"]] == status
            assert [tuple(r["pos"]) for r in s["rocks"]] == [tuple(p) for p in cfg.rock_pos[:k]]
            assert env._generate_legal() == O.rock_generate_legal(cfg, x, y, status)
            if done:
                with pytest.raises(AssertionError):       # rock.py:126
                    env.step(0)
                break


def test_rock_set_state_roundtrip_and_compute_prob(backend):
    env = gp.make("Rock-v0", board_size=7, num_rocks=8, device=backend, seed=SEED)
    cfg = O.RockCfg(7, 8)
    env.reset()
    s = env._get_init_state()
    assert s["agent_pos"] == Coord(*cfg.start) and len(s["rocks"]) == 8
    s["agent_pos"] = Coord(2, 0)                             # on rock 0
    s["rocks"][0]["status"] = 1
    env._set_state(s)
    ob, rw, done, info = env.step(4)                         # sample a good rock
    assert (ob, rw, done) == (0, 10, False) and info["state"]["rocks"][0]["status"] == 0
    ob, rw, done, info = env.step(4)                         # nothing left: -100 and terminal (rock.py:193)
    assert (rw, done) == (-100, True)
    # _compute_prob: rock.py:250-264
    env._set_state(s)
    st = env._get_init_state()
    st["agent_pos"] = Coord(2, 0)
    for a in range(13):
        for ob in range(3):
            exp = O.rock_compute_prob(cfg, a, 2, 0, [r["status"] for r in st["rocks"]], ob)
            assert env._compute_prob(a, st, ob) == exp


def test_rock15_dangling_cell_raises_like_the_reference(backend):
    env = gp.make("Rock-v0", board_size=15, num_rocks=15, device=backend, seed=SEED)
    env.reset()
    s = env._get_init_state()
    s["agent_pos"] = Coord(12, 2)
    env._set_state(s)
    with pytest.raises(IndexError):                          # rock.py:162 on grid id 15
        env.step(4)


def test_stochastic_rock_episode_matches_oracle(backend):
    env = gp.make("StochasticRock-v0", board_size=7, num_rocks=8, device=backend, seed=SEED)
    cfg = O.RockCfg(7, 8, stochastic=True)
    env.reset()
    x, y, status, _ = O.rock_reset(cfg, words(1, philox.DOMAIN_RESET, 16))
    rs = np.random.RandomState(3)
    ctr = 1
    for _ in range(80):
        a = int(rs.randint(0, 13))
        ctr += 1
        ob, rw, done, _ = env.step(a)
        x, y, status, eob, erw, edone, _ = O.rock_step(cfg, x, y, status, a, words(ctr, philox.DOMAIN_STEP))
        assert (ob, rw, done) == (eob, erw, edone)
        if done:
            break


def test_tag_episode_matches_oracle(backend):
    for n_opp in (1, 2):
        env = gp.make("Tag-v0", num_opponents=n_opp, device=backend, seed=SEED)
        ob = env.reset()
        ax, ay, opps, num_opp, eob = O.tag_reset(n_opp, words(1, philox.DOMAIN_RESET))
        assert ob == eob
        rs = np.random.RandomState(n_opp)
        ctr = 1
        for _ in range(200):
            a = int(rs.randint(0, 5))
            ctr += 1
            ob, rw, done, info = env.step(a)
            ax, ay, opps, num_opp, eob, erw, edone = O.tag_step(ax, ay, opps, num_opp, a, words(ctr, philox.DOMAIN_STEP))
            assert (ob, rw, done) == (eob, erw, edone) and isinstance(rw, float)
            s = info["state"]
            assert tuple(s.agent_pos) == (ax, ay) and [tuple(o) for o in s.opponent_pos] == [tuple(o) for o in opps]
            assert s.num_opp == num_opp
            if done:
                break


def test_tiger_episode_matches_oracle(backend):
    env = gp.make("Tiger-v0", device=backend, seed=SEED)
    for ep in range(6):
        ctr0 = env._step_ctr
        assert env.reset() == 2
        state, _ = O.tiger_reset(words(ctr0 + 1, philox.DOMAIN_RESET))
        rs = np.random.RandomState(ep)
        ctr = ctr0 + 1
        for _ in range(30):
            a = int(rs.choice([2, 2, 2, 0, 1]))
            ctr += 1
            ob, rw, done, info = env.step(a)
            state, eob, erw, edone = O.tiger_step(state, a, words(ctr, philox.DOMAIN_STEP))
            assert (ob, rw, done, info["state"]) == (eob, erw, edone, state)
            if done:
                assert ob == state                                  # tiger.py:81-83
                break


def test_network_episode_matches_oracle(backend):
    env = gp.make("Network-v0", device=backend, seed=SEED)
    assert env.reset() == 0                                         # network.py:69: Obs.OFF, not NULL
    nb = O.network_neighbours(10, 3)
    assert env.neighbours == nb
    state, _ = O.network_reset(10)
    rs = np.random.RandomState(4)
    ctr = 1
    for _ in range(120):
        a = int(rs.randint(0, 21))
        ctr += 1
        ob, rw, done, info = env.step(a)
        w = philox.network_draws(SEED, np.arange(1), ctr, 10)[0]
        state, eob, tenths, _ = O.network_step(state, a, lambda slot: int(w[slot]), nb)
        assert (ob, done) == (eob, False)
        assert rw == tenths / 10.0                                  # the reference's double, exactly
        assert list(info["state"]) == state


def test_battleship_episode_matches_oracle(backend):
    for size in [(5, 5), (10, 10)]:
        env = gp.make("Battleship-v0", board_size=size, device=backend, seed=SEED, reset_mode="rejection")
        assert env.reset() == 0
        board, ok = O.battleship_reset_rejection(size[0], size[1], 3, words(1, philox.DOMAIN_RESET, 512))
        assert ok
        order = np.random.RandomState(7).permutation(size[0] * size[1])
        order = np.concatenate([order[:3], order[:2], order[3:]])           # a few repeated shots: -10
        for a in order:
            ob, rw, done, info = env.step(int(a))
            eob, erw, edone = O.battleship_step(board, int(a))
            assert (ob, rw, done) == (eob, erw, edone) and isinstance(rw, int)
            assert info["state"].total_remaining == board.total_remaining
            if done:
                with pytest.raises(AssertionError):                         # battleship.py:93
                    env.step(0)
                break
        assert done


def test_batched_step_protocol(backend):
    B = 64
    env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=B, device=backend, seed=SEED)
    with pytest.raises(AssertionError):
        env.step(torch.zeros(B, dtype=torch.int32))
    ob = env.reset()
    assert ob.shape == (B,) and ob.dtype == torch.int32 and not ob.any()
    ob, rw, done, info = env.step(torch.ones(B, dtype=torch.int64))
    assert ob.dtype == torch.int32 and rw.dtype == torch.float32 and done.dtype == torch.bool
    assert info["state"].shape == (B,) and info["flags"].shape == (B,)
    x, y, st, dn = env.unpack(info["state"])
    assert (x == 1).all() and (y == 5).all() and not dn.any()          # config[11] start (0,5), one step EAST
    with pytest.raises(ValueError):
        env.step(torch.ones(B + 1, dtype=torch.int32))
    arr = env.to_array_form(info["state"])                            # rock.py:205-210
    assert arr.shape == (B, 12) and (arr[:, 0] == 11 * 5 + 1).all()


# ------------------------------------------------------------ gym / gymnasium registration ---
def _stub_gym(name, version):
    """A stand-in for the parts of gym / gymnasium that registration touches (neither package is in the image): Env,
    envs.registration.register / make, and -- for the new API -- the behaviour that broke old-API envs: make() wraps the
    env in a checker that calls reset(seed=..., options=...) and insists on (ob, info) and five values from step()."""
    import types
    mod = types.ModuleType(name)
    mod.__version__ = version

    class Env(object):
        metadata = {}
    mod.Env = Env
    reg = types.ModuleType(name + ".envs.registration")
    reg.specs = {}

    def register(id, entry_point=None, **kw):
        reg.specs[id] = (entry_point, kw)
    reg.register = register
    new_api = name == "gymnasium" or tuple(int(p) for p in version.split(".")[:2]) >= (0, 26)

    class OrderEnforcing(Env):
        def __init__(self, env):
            assert isinstance(env, Env), "gym.make wraps only gym.Env instances"
            self.env = env

        def reset(self, *, seed=None, options=None):
            out = self.env.reset(seed=seed, options=options)
            assert isinstance(out, tuple) and len(out) == 2 and isinstance(out[1], dict)
            return out

        def step(self, a):
            out = self.env.step(a)
            assert len(out) == 5
            return out

    def make(id, **kw):
        entry, spec_kw = reg.specs[id]
        if callable(entry):
            env = entry(**kw)
        else:
            import importlib
            m, _, cls = entry.partition(":")
            env = getattr(importlib.import_module(m), cls)(**kw)
        if new_api:
            assert spec_kw.get("disable_env_checker") is True          # the passive checker would reject batched outputs
            return OrderEnforcing(env)
        return env
    reg.make = make
    envs = types.ModuleType(name + ".envs")
    envs.registration = reg
    mod.envs = envs
    return {name: mod, name + ".envs": envs, name + ".envs.registration": reg}


@pytest.mark.parametrize("pkg,version", [("gymnasium", "0.29.1"), ("gym", "0.26.2"), ("gym", "0.21.0")])
def test_registration_with_old_and_new_gym_apis(backend, monkeypatch, pkg, version):
    """the envs speak the reference's old-gym protocol; gym >= 0.26 / gymnasium call reset(seed=, options=) and
    expect (ob, info) and a 5-tuple.  They get an adapter; old gym gets the classes themselves."""
    import sys
    from gym_pomdp_b200 import registration
    stubs = _stub_gym(pkg, version)
    for k in ("gym", "gym.envs", "gym.envs.registration", "gymnasium", "gymnasium.envs", "gymnasium.envs.registration"):
        monkeypatch.delitem(sys.modules, k, raising=False)
    for k, v in stubs.items():
        monkeypatch.setitem(sys.modules, k, v)
    registration._adapter_class.cache_clear()
    assert registration.register_with_gym() == pkg
    reg = stubs[pkg + ".envs.registration"]
    assert sorted(reg.specs) == sorted(gp.registry)
    new_api = registration.uses_new_api(pkg)
    assert new_api == (pkg == "gymnasium" or version.startswith("0.26"))
    env = reg.make("Tiger-v0", device=backend, seed=SEED)
    if new_api:
        ob, info = env.reset(seed=11, options=None)
        assert ob == 2 and info == {}
        ob, rw, terminated, truncated, info = env.step(2)
        assert ob in (0, 1) and rw == -1 and terminated is False and truncated is False and "state" in info
        inner = env.env
        assert isinstance(inner, stubs[pkg].Env) and inner._generate_legal() == [0, 1, 2] and inner._discount == .95
        assert inner.env._seed == 11                                   # reset(seed=) reached env.seed()
        benv = reg.make("Tiger-v0", device=backend, batch_size=8).env
        ob, info = benv.reset(seed=3)
        out = benv.step(torch.full((8,), 2, dtype=torch.int32))
        assert len(out) == 5 and out[3].shape == out[2].shape and not out[3].any()
    else:
        assert isinstance(env, gp.envs.TigerEnv) and env.reset() == 2 and len(env.step(2)) == 4
    # the env classes subclass gym.Env only where the old protocol is gym's own
    from gym_pomdp_b200.envs import base
    assert (base._env_base() is stubs[pkg].Env) == (pkg == "gym" and not new_api)
    registration._adapter_class.cache_clear()
    # reset() itself tolerates the new keywords (a caller that skips gym.make)
    plain = gp.make("Tiger-v0", device=backend)
    assert plain.reset(seed=5, options={}) == 2 and plain._seed == 5


def test_default_seed_is_fresh_and_explicit_seed_reproduces(backend):
    """the reference draws from numpy's unseeded global RNG, so two default-constructed envs are independent;
    here the default seed is a fresh 64-bit key, an explicit one reproduces."""
    a = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=256, device=backend)
    b = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=256, device=backend)
    assert a._seed != b._seed and not torch.equal(a.reset() * 0 + a.state, b.reset() * 0 + b.state)
    c = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=256, device=backend, seed=5)
    d = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=256, device=backend, seed=5)
    c.reset(), d.reset()
    assert torch.equal(c.state, d.state)
    assert a.seed(9) == [9] and a._seed == 9 and a.seed() == [None] and a._seed != 9
