"""Packed-state kernels vs the fixtures recorded from the UNMODIFIED reference.

Every test runs twice: on the g++ host build of the kernel functors (CPU suite) and,
marked ``gpu``, on the real CUDA library through the C ABI.  The fixtures' draw tables are
Philox words of (seed 0x5EED, env = case index, step = fixture stream), which is exactly
what the kernels regenerate in registers -- so stochastic transitions compare ELEMENT-WISE
with what the reference did under the same uniforms (coupling rules: oracle/ref_shim.py).
Bar: bit-exact (integers) and exact float32 equality for rewards.
"""
import numpy as np
import pytest
import torch

import gym_pomdp_b200 as gp
from gym_pomdp_b200 import _lib
from oracle import philox

from backends import GOLDEN_SEED, backend, philox_unmodified  # noqa: F401

ROCKS = ["7_8", "11_11", "15_15", "7_7", "4_3", "stoch_7_8", "stoch_11_11"]


def t(a, dev, dtype=torch.int32):
    return torch.as_tensor(np.asarray(a), device=dev).to(dtype)


@pytest.mark.parametrize("tag", ROCKS)
def test_rock_step_vs_reference(golden, backend, tag):
    g = golden("rock_" + tag)
    n, k, stoch = int(g["n"]), int(g["k"]), bool(g["stochastic"])
    env = gp.make("StochasticRock-v0" if stoch else "Rock-v0", board_size=n, num_rocks=k, batch_size=len(g["x"]),
                  device=backend, seed=GOLDEN_SEED)
    assert env.state_words == (1 if k <= 11 else 2)
    # static maps built by the C library == the reference's grid / efficiency table
    grid = env._grid_map.reshape(16, 16)[:n, :n].T      # [y, x] -> [x, y]
    assert np.array_equal(grid, g["grid"])
    assert np.array_equal(env.grid.board, g["grid"]) and env.grid.board.dtype == np.int8       # coord.py:51-52, rock.py:110-111
    assert env.grid[(0, 0)] == g["grid"][0, 0] and env.grid[(n, 0)] is None                    # coord.py:45-49
    assert np.array_equal(env._eff_T[: 2 * (n - 1) + 1], np.ceil(g["eff"] * 2.0 ** 32).astype(np.int64))
    assert [tuple(p) for p in env._rock_pos] == [tuple(p) for p in g["rock_pos"][:k]]
    state = env.pack(g["x"], g["y"], g["status"])
    ns, ob, rw, fl = env.simulate(state, t(g["action"], backend), step_ctr=11)
    x2, y2, st2, done = (v.cpu().numpy() for v in env.unpack(ns))
    ob, rw, fl = ob.cpu().numpy(), rw.cpu().numpy(), fl.cpu().numpy()
    ok = philox_unmodified(g["draws"], 11, philox.DOMAIN_STEP) & ~g["raised"]
    assert ok.sum() > 0.7 * len(ok)
    assert np.array_equal(x2[ok], g["x2"][ok]) and np.array_equal(y2[ok], g["y2"][ok])
    assert np.array_equal(st2[ok], g["status2"][ok])
    assert np.array_equal(ob[ok], g["obs"][ok])
    assert np.array_equal(rw[ok], g["reward"][ok].astype(np.float32))
    assert np.array_equal(done[ok], g["done"][ok])
    assert np.array_equal((fl[ok] & 1).astype(bool), g["done"][ok]) and not (fl[ok] & ~1).any()
    # where the reference raised IndexError (dangling grid id): flagged, treated as "no rock"
    r = g["raised"]
    if r.any():
        assert ((fl[r] & _lib.FLAG_BAD_STATE) != 0).all()
    # NEXT rows: batched observation likelihood on the post state
    for o in range(3):
        p = env._compute_prob(t(g["action"], backend), ns, torch.full((len(ok),), o, device=backend))
        assert np.array_equal(p.cpu().numpy()[ok], g["prob"][ok, o])
    legal = env._generate_legal(ns).cpu().numpy()
    for i in np.nonzero(ok)[0][:600]:
        exp = sorted(set(a for a in g["legal"][i].tolist() if a >= 0))
        if exp:
            assert np.nonzero(legal[i])[0].tolist() == exp


@pytest.mark.parametrize("tag", ROCKS)
def test_rock_reset_vs_reference(golden, backend, tag):
    g = golden("rock_" + tag)
    n, k = int(g["n"]), int(g["k"])
    M = len(g["reset_draws"])
    env = gp.make("Rock-v0", board_size=n, num_rocks=k, batch_size=M, device=backend, seed=GOLDEN_SEED)
    state, obs = env.init_states(M, step_ctr=0)
    x, y, st, done = (v.cpu().numpy() for v in env.unpack(state))
    ok = philox_unmodified(g["reset_draws"], 0, philox.DOMAIN_RESET)
    assert ok.sum() >= M - 10
    assert (x == g["start"][0]).all() and (y == g["start"][1]).all() and not done.any()
    assert np.array_equal(st[ok], g["reset_status"][ok])
    assert (obs.cpu().numpy() == 0).all()


@pytest.mark.parametrize("tag", ["1opp", "2opp"])
def test_tag_vs_reference(golden, backend, tag):
    g = golden("tag_" + tag)
    n_opp = int(g["n_opp"])
    N = len(g["agent"])
    env = gp.make("Tag-v0", num_opponents=n_opp, batch_size=N, device=backend, seed=GOLDEN_SEED)
    state = env.pack(g["agent"], g["opp"], num_opp=g["num_opp"])
    ns, ob, rw, fl = env.simulate(state, t(g["action"], backend), step_ctr=12)
    agent2, opp2, nop2, done = (v.cpu().numpy() for v in env.unpack(ns))
    ok = philox_unmodified(g["draws"], 12, philox.DOMAIN_STEP)
    assert ok.sum() >= N - 200
    assert np.array_equal(agent2[ok], g["agent2"][ok])
    assert np.array_equal(opp2[ok], g["opp2"][ok])
    assert np.array_equal(nop2[ok], g["num_opp2"][ok])
    assert np.array_equal(ob.cpu().numpy()[ok], g["obs"][ok])
    assert np.array_equal(rw.cpu().numpy()[ok], g["reward"][ok].astype(np.float32))
    assert np.array_equal(done[ok], g["done"][ok])
    assert np.array_equal(fl.cpu().numpy()[ok], g["done"][ok].astype(np.int32))
    for o in (0, 28, 29):
        p = env._compute_prob(t(g["action"], backend), ns, torch.full((N,), o, device=backend))
        assert np.array_equal(p.cpu().numpy()[ok], g["prob"][ok, o])
    # reset
    M = len(g["reset_draws"])
    st, rob = env.init_states(M, step_ctr=0)
    ragent, ropp, rnop, _ = (v.cpu().numpy() for v in env.unpack(st))
    rok = philox_unmodified(g["reset_draws"], 0, philox.DOMAIN_RESET)
    assert np.array_equal(ragent[rok], g["reset_agent"][rok]) and np.array_equal(ropp[rok], g["reset_opp"][rok])
    assert np.array_equal(rob.cpu().numpy()[rok], g["reset_obs"][rok]) and (rnop == n_opp).all()


def test_tiger_vs_reference(golden, backend):
    g = golden("tiger")
    N = len(g["state"])
    env = gp.make("Tiger-v0", batch_size=N, device=backend, seed=GOLDEN_SEED)
    ns, ob, rw, fl = env.simulate(env.pack(g["state"]), t(g["action"], backend), step_ctr=13)
    s2, done = (v.cpu().numpy() for v in env.unpack(ns))
    ok = philox_unmodified(g["draws"], 13, philox.DOMAIN_STEP)
    assert np.array_equal(s2[ok], g["state2"][ok])
    assert np.array_equal(ob.cpu().numpy()[ok], g["obs"][ok])
    assert np.array_equal(rw.cpu().numpy()[ok], g["reward"][ok].astype(np.float32))
    assert np.array_equal(done[ok], g["done"][ok])
    for o in range(3):
        p = env._compute_prob(t(g["action"], backend), ns, torch.full((N,), o, device=backend))
        assert np.array_equal(p.cpu().numpy()[ok], g["prob"][ok, o])
    M = len(g["reset_draws"])
    st, rob = env.init_states(M, step_ctr=0)
    assert np.array_equal(env.unpack(st)[0].cpu().numpy(), g["reset_state"])
    assert np.array_equal(rob.cpu().numpy(), g["reset_obs"])


@pytest.mark.parametrize("tag", ["3legs10", "3legs7", "ring10", "3legs19"])
def test_network_vs_reference(golden, backend, tag):
    g = golden("network_" + tag)
    n, N = int(g["n"]), len(g["state"])
    env = gp.make("Network-v0", n_machines=n, problem_type=int(g["problem_type"]), batch_size=N, device=backend,
                  seed=GOLDEN_SEED)
    assert env.neighbours == [[j for j in row if j >= 0] for row in g["neighbours"].tolist()]
    state = t(g["state"], backend)
    ns, ob, rw, fl = env.simulate(state, t(g["action"], backend), step_ctr=14)
    assert np.array_equal(ns.cpu().numpy().astype(np.int64), g["state2"])
    assert np.array_equal(ob.cpu().numpy(), g["obs"])
    assert np.array_equal(rw.cpu().numpy(), g["reward"].astype(np.float32))   # float32(double reward), exactly
    assert not fl.cpu().numpy().any()
    for o in range(3):
        p = env._compute_prob(t(g["action"], backend), ns, torch.full((N,), o, device=backend))
        assert np.array_equal(p.cpu().numpy(), g["prob"][:, o])
    st, rob = env.init_states(4)
    assert (st.cpu().numpy() == int(g["reset_state"])).all() and (rob.cpu().numpy() == int(g["reset_obs"])).all()


@pytest.mark.parametrize("tag", ["10x10", "5x5"])
def test_battleship_vs_reference(golden, backend, tag):
    g = golden("battleship_" + tag)
    xs, ys, max_len = int(g["x_size"]), int(g["y_size"]), int(g["max_len"])
    B = len(g["occupied"])
    # reset, rejection flavour: the reference's own loop under the same draws -> identical boards
    env = gp.make("Battleship-v0", board_size=(xs, ys), max_len=max_len, batch_size=B, device=backend,
                  seed=GOLDEN_SEED, reset_mode="rejection")
    st, rob = env.init_states(B, step_ctr=0)
    occ, vis, rem, done = (v.cpu().numpy() for v in env.unpack(st))
    assert np.array_equal(occ, g["occupied"]) and not vis.any() and (rem == 5).all() and not done.any()
    assert (rob.cpu().numpy() == 0).all() and not env.reset_flags.cpu().numpy().any()
    # steps from the fixture's synthetic (visited, total_remaining) states
    state = env.pack(g["occupied"], g["visited_in"], total_remaining=g["remaining"][:, 0])
    alive = np.ones(B, bool)
    for s in range(g["action"].shape[1]):
        a = g["action"][:, s]
        alive &= a >= 0
        ns, ob, rw, fl = env.simulate(state, t(np.where(a >= 0, a, 0), backend))
        _, _, rem2, done2 = (v.cpu().numpy() for v in env.unpack(ns))
        assert np.array_equal(ob.cpu().numpy()[alive], g["obs"][alive, s])
        assert np.array_equal(rw.cpu().numpy()[alive], g["reward"][alive, s].astype(np.float32))
        assert np.array_equal(done2[alive], g["done"][alive, s])
        assert np.array_equal(rem2[alive], g["remaining"][alive, s + 1])
        for o in range(2):
            p = env._compute_prob(t(np.where(a >= 0, a, 0), backend), ns, torch.full((B,), o, device=backend))
            assert np.array_equal(p.cpu().numpy()[alive], g["prob"][alive, s, o])
        assert np.array_equal(env._generate_legal(ns).sum(dim=1).cpu().numpy()[alive], g["legal_count"][alive, s])
        state = ns
    # stepping a finished board is flagged, not executed
    fin = ~alive
    if fin.any():
        ns, ob, rw, fl = env.simulate(state, t(np.zeros(B), backend))
        assert ((fl.cpu().numpy()[fin] & _lib.FLAG_STEPPED_DONE) != 0).all()
        assert np.array_equal(ns.cpu().numpy()[fin], state.cpu().numpy()[fin])


@pytest.mark.parametrize("tag", ["10x10", "5x5"])
def test_battleship_scan_reset_accepts_exactly_the_reference_set(golden, backend, tag):
    """The warp-scan reset must choose among exactly the placements the reference's
    collision() accepts.  Checked through the kernel's own output: ship 1 always lands on a
    fixture-valid first placement, and given it, ship 2 on a fixture-valid second one."""
    g = golden("battleship_" + tag)
    xs, ys, max_len = int(g["x_size"]), int(g["y_size"]), int(g["max_len"])
    B = 4096
    env = gp.make("Battleship-v0", board_size=(xs, ys), max_len=max_len, batch_size=B, device=backend, seed=1234)
    st, _ = env.init_states(B, step_ctr=5)
    occ, vis, rem, done = (v.cpu().numpy() for v in env.unpack(st))
    assert (occ.reshape(B, -1).sum(1) == 5).all() and (rem == 5).all() and not vis.any() and not done.any()
    assert not env.reset_flags.cpu().numpy().any()
    # every produced board must be reachable by the reference: decompose into a 3-ship and a 2-ship
    valid_first = set(np.nonzero(g["valid_first"])[0].tolist())
    from oracle import pomdp_oracle as O
    seen_first = set()
    for b in range(0, B, 16):
        found = False
        for c in valid_first:
            x, y = O.grid_get_coord(xs, c >> 2)
            dx, dy = O.COMPASS[c & 3]
            cells = [(x + i * dx, y + i * dy) for i in range(3)]
            if all(occ[b, cx, cy] for cx, cy in cells):
                board = O.ShipBoard(xs, ys)
                O.ship_mark(board, x, y, c & 3, 3)
                rest = occ[b].copy()
                for cx, cy in cells:
                    rest[cx, cy] = False
                for c2 in O.battleship_valid_placements(board, 2):
                    x2, y2 = O.grid_get_coord(xs, c2 >> 2)
                    d2 = O.COMPASS[c2 & 3]
                    if rest[x2, y2] and rest[x2 + d2[0], y2 + d2[1]] and rest.sum() == 2:
                        found = True
                        seen_first.add(c)
                        break
            if found:
                break
        assert found, b
    assert len(seen_first) > 10
