"""Uniform-legal policy and fused T-step rollouts (the loops at
rock.py:563-572 and tag.py:310-316).

Chain of evidence:
  1. tests/golden/rollouts.npz holds episodes the UNMODIFIED reference played itself
     (``np.random.choice(env._generate_legal())`` + ``env.step`` + ``r += rw * discount``), its
     draws scripted from Philox words (oracle/gen_rollouts.py);
  2. the Python and C oracles reproduce those episodes (CPU, pins the oracle);
  3. the fused rollout kernel reproduces them too -- float64 returns compared with ``==`` -- and
     matches the C oracle on large random batches;
  4. the fused kernel equals, draw for draw, T launches of the policy kernel + the step kernel.
"""
import numpy as np
import pytest
import torch

import gym_pomdp_b200 as gp
from gym_pomdp_b200 import _lib
from oracle import c_oracle as C
from oracle import philox
from oracle import pomdp_oracle as O

from backends import backend  # noqa: F401

SEED = 0x5EED
ROCKS = ["rock_7_8", "rock_11_11", "rock_15_15", "srock_7_8"]


def draws_for(e, ctr0):
    def draws(t):
        w = philox.draw_slots(SEED, np.array([e]), ctr0 + t, philox.DOMAIN_POLICY, 1)[0, 0]
        sw = philox.draw_slots(SEED, np.array([e]), ctr0 + t, philox.DOMAIN_STEP, 12)[0]
        return int(w), (lambda slot: int(sw[slot]))
    return draws


# ------------------------------------------------------------------ oracle vs reference ---
@pytest.mark.parametrize("tag", ROCKS)
def test_oracles_replay_reference_rock_rollouts(golden, tag):
    g = golden("rollouts")
    n, k, stoch, T = (int(v) for v in g[tag + "_cfg"])
    cfg = O.RockCfg(n, k, bool(stoch))
    gamma = float(g[tag + "_discount"])
    M = len(g[tag + "_ret"])
    for e in range(0, M, 3):
        st = {"x": int(g[tag + "_x0"][e]), "y": int(g[tag + "_y0"][e]), "s": g[tag + "_st0"][e].tolist(), "done": False}

        def step(a, draw):
            st["x"], st["y"], st["s"], ob, rw, st["done"], err = O.rock_step(cfg, st["x"], st["y"], st["s"], a, draw)
            return rw, st["done"]
        ret, steps, acts = O.rollout(lambda: O.rock_generate_legal(cfg, st["x"], st["y"], st["s"]), step, lambda: st["done"],
                                     T, gamma, draws_for(e, int(g["first_ctr"])))
        assert ret == g[tag + "_ret"][e] and steps == g[tag + "_steps"][e]
        assert acts == [a for a in g[tag + "_acts"][e].tolist() if a >= 0]
        assert (st["x"], st["y"], st["s"], st["done"]) == (g[tag + "_x1"][e], g[tag + "_y1"][e], g[tag + "_st1"][e].tolist(),
                                                           bool(g[tag + "_done"][e]))
    x1, y1, s1, ret, steps, done, err = C.rock_rollout(n, k, stoch, 0.8, g[tag + "_x0"], g[tag + "_y0"], g[tag + "_st0"], SEED, 0,
                                                       int(g["first_ctr"]), T, gamma)
    assert np.array_equal(ret, g[tag + "_ret"]) and np.array_equal(steps, g[tag + "_steps"])
    assert np.array_equal(x1, g[tag + "_x1"]) and np.array_equal(y1, g[tag + "_y1"]) and np.array_equal(s1, g[tag + "_st1"])
    assert np.array_equal(done, g[tag + "_done"]) and not err.any()


def test_c_oracle_replays_reference_rollouts_other_envs(golden):
    g = golden("rollouts")
    c0 = int(g["first_ctr"])
    for tag in ("tag_1opp", "tag_2opp"):
        n_opp, T = (int(v) for v in g[tag + "_cfg"])
        M = len(g[tag + "_ret"])
        a1, o1, n1, ret, steps, done = C.tag_rollout(n_opp, 0.8, g[tag + "_agent0"], g[tag + "_opp0"], np.full(M, n_opp), SEED, 0,
                                                     c0, T, float(g[tag + "_discount"]))
        assert np.array_equal(ret, g[tag + "_ret"]) and np.array_equal(steps, g[tag + "_steps"])
        assert np.array_equal(a1, g[tag + "_agent1"]) and np.array_equal(o1, g[tag + "_opp1"].reshape(M, n_opp))
        assert np.array_equal(n1, g[tag + "_nopp1"]) and np.array_equal(done, g[tag + "_done"])
    s1, ret, steps, done = C.tiger_rollout(0.85, g["tiger_s0"], SEED, 0, c0, int(g["tiger_cfg"][0]), float(g["tiger_discount"]))
    assert np.array_equal(ret, g["tiger_ret"]) and np.array_equal(steps, g["tiger_steps"]) and np.array_equal(s1, g["tiger_s1"])
    assert np.array_equal(done, g["tiger_done"])
    n, ptype, T = (int(v) for v in g["network_cfg"])
    M = len(g["network_ret"])
    m1, ret, steps = C.network_rollout(n, ptype, np.ones((M, n), np.int8), SEED, 0, c0, T, float(g["network_discount"]))
    assert np.array_equal(ret, g["network_ret"]) and (steps == T).all()
    assert np.array_equal((m1.astype(np.int64) << np.arange(n)).sum(1), g["network_s1"])
    for tag in ("ship_5x5", "ship_10x10"):
        xs, ys, max_len, T = (int(v) for v in g[tag + "_cfg"])
        M = len(g[tag + "_ret"])
        v1, r1, ret, steps, done = C.battleship_rollout(xs, ys, g[tag + "_occ"], np.zeros((M, xs, ys), np.uint8), np.full(M, 5), SEED,
                                                        0, c0, T, float(g[tag + "_discount"]))
        assert np.array_equal(ret, g[tag + "_ret"]) and np.array_equal(steps, g[tag + "_steps"])
        assert np.array_equal(v1, g[tag + "_vis1"]) and np.array_equal(r1, g[tag + "_rem1"]) and np.array_equal(done, g[tag + "_done"])


# ------------------------------------------------------------------ kernels vs reference ---
@pytest.mark.parametrize("tag", ROCKS)
def test_rock_rollout_kernel_vs_reference(golden, backend, tag):
    g = golden("rollouts")
    n, k, stoch, T = (int(v) for v in g[tag + "_cfg"])
    M = len(g[tag + "_ret"])
    env = gp.make("StochasticRock-v0" if stoch else "Rock-v0", board_size=n, num_rocks=k, batch_size=M, device=backend, seed=SEED)
    state, _ = env.init_states(M, step_ctr=int(g["reset_ctr"]))
    x0, y0, st0, _ = (v.cpu().numpy() for v in env.unpack(state))
    assert np.array_equal(x0, g[tag + "_x0"]) and np.array_equal(y0, g[tag + "_y0"]) and np.array_equal(st0, g[tag + "_st0"])
    # the first action of every episode, through the policy kernel
    a0 = env.sample_legal_actions(state, step_ctr=int(g["first_ctr"])).cpu().numpy()
    assert np.array_equal(a0, g[tag + "_acts"][:, 0])
    final, ret, steps, flags = env.rollout(state, max_steps=T, step_ctr=int(g["first_ctr"]))
    assert np.array_equal(ret.cpu().numpy(), g[tag + "_ret"])              # float64, exactly the Python loop's value
    assert np.array_equal(steps.cpu().numpy(), g[tag + "_steps"])
    x1, y1, st1, done = (v.cpu().numpy() for v in env.unpack(final))
    assert np.array_equal(x1, g[tag + "_x1"]) and np.array_equal(y1, g[tag + "_y1"]) and np.array_equal(st1, g[tag + "_st1"])
    assert np.array_equal(done, g[tag + "_done"]) and np.array_equal(flags.cpu().numpy(), g[tag + "_done"].astype(np.int32))


def test_other_rollout_kernels_vs_reference(golden, backend):
    g = golden("rollouts")
    c0, cr = int(g["first_ctr"]), int(g["reset_ctr"])
    for tag in ("tag_1opp", "tag_2opp"):
        n_opp, T = (int(v) for v in g[tag + "_cfg"])
        M = len(g[tag + "_ret"])
        env = gp.make("Tag-v0", num_opponents=n_opp, batch_size=M, device=backend, seed=SEED)
        state, _ = env.init_states(M, step_ctr=cr)
        final, ret, steps, flags = env.rollout(state, max_steps=T, step_ctr=c0)
        a1, o1, n1, done = (v.cpu().numpy() for v in env.unpack(final))
        assert np.array_equal(ret.cpu().numpy(), g[tag + "_ret"]) and np.array_equal(steps.cpu().numpy(), g[tag + "_steps"])
        assert np.array_equal(a1, g[tag + "_agent1"]) and np.array_equal(o1, g[tag + "_opp1"].reshape(M, n_opp))
        assert np.array_equal(n1, g[tag + "_nopp1"]) and np.array_equal(done, g[tag + "_done"])
        assert np.array_equal(env.sample_legal_actions(state, step_ctr=c0).cpu().numpy(), g[tag + "_acts"][:, 0])
    M = len(g["tiger_ret"])
    env = gp.make("Tiger-v0", batch_size=M, device=backend, seed=SEED)
    state, _ = env.init_states(M, step_ctr=cr)
    final, ret, steps, flags = env.rollout(state, max_steps=int(g["tiger_cfg"][0]), step_ctr=c0)
    assert np.array_equal(ret.cpu().numpy(), g["tiger_ret"]) and np.array_equal(steps.cpu().numpy(), g["tiger_steps"])
    s1, done = (v.cpu().numpy() for v in env.unpack(final))
    assert np.array_equal(s1, g["tiger_s1"]) and np.array_equal(done, g["tiger_done"])
    n, ptype, T = (int(v) for v in g["network_cfg"])
    M = len(g["network_ret"])
    env = gp.make("Network-v0", n_machines=n, problem_type=ptype, batch_size=M, device=backend, seed=SEED)
    state, _ = env.init_states(M)
    final, ret, steps, flags = env.rollout(state, max_steps=T, step_ctr=c0)
    assert np.array_equal(ret.cpu().numpy(), g["network_ret"])            # sums of s - 0.1 / s - 2.5 doubles, bit for bit
    assert np.array_equal(final.cpu().numpy().astype(np.int64), g["network_s1"]) and (steps == T).all() and not flags.any()
    for tag in ("ship_5x5", "ship_10x10"):
        xs, ys, max_len, T = (int(v) for v in g[tag + "_cfg"])
        M = len(g[tag + "_ret"])
        env = gp.make("Battleship-v0", board_size=(xs, ys), batch_size=M, device=backend, seed=SEED, reset_mode="rejection")
        state, _ = env.init_states(M, step_ctr=cr)
        occ, vis, rem, done = (v.cpu().numpy() for v in env.unpack(state))
        assert np.array_equal(occ.reshape(M, xs, ys), g[tag + "_occ"])
        final, ret, steps, flags = env.rollout(state, max_steps=T, step_ctr=c0)
        _, v1, r1, done = (v.cpu().numpy() for v in env.unpack(final))
        assert np.array_equal(ret.cpu().numpy(), g[tag + "_ret"]) and np.array_equal(steps.cpu().numpy(), g[tag + "_steps"])
        assert np.array_equal(v1.reshape(M, xs, ys), g[tag + "_vis1"]) and np.array_equal(r1, g[tag + "_rem1"])
        assert np.array_equal(done, g[tag + "_done"])
        assert np.array_equal(env.sample_legal_actions(state, step_ctr=c0).cpu().numpy(), g[tag + "_acts"][:, 0])


# --------------------------------------------------------- kernels vs C oracle, large batches ---
def big(backend):
    return (1 << 16) if backend.startswith("cuda") else (1 << 11)


@pytest.mark.parametrize("board,k,stoch", [(11, 11, False), (15, 15, False), (7, 8, True)])
def test_rock_rollout_kernel_vs_c_oracle(backend, board, k, stoch):
    N, T, goff = big(backend), 64, 4 * 12345
    env = gp.make("StochasticRock-v0" if stoch else "Rock-v0", board_size=board, num_rocks=k, batch_size=N, device=backend,
                  seed=SEED, global_offset=goff)
    rs = np.random.RandomState(board)
    x, y, status = rs.randint(0, board, N), rs.randint(0, board, N), rs.randint(-1, 2, (N, k))
    if board == 15:
        x[:64], y[:64] = 12, 2                                  # start on the dangling cell: SAMPLE is never offered
    state = env.pack(x, y, status)
    final, ret, steps, flags = env.rollout(state, max_steps=T, step_ctr=9, discount=0.95)
    ex, ey, est, eret, esteps, edone, err = C.rock_rollout(board, k, stoch, 0.8, x, y, status, SEED, goff, 9, T, 0.95)
    x1, y1, st1, done = (v.cpu().numpy() for v in env.unpack(final))
    assert np.array_equal(ret.cpu().numpy(), eret) and np.array_equal(steps.cpu().numpy(), esteps)
    assert np.array_equal(x1, ex) and np.array_equal(y1, ey) and np.array_equal(st1, est.astype(np.int32))
    assert np.array_equal(done, edone) and not err.any() and np.array_equal(flags.cpu().numpy(), edone.astype(np.int32))
    assert edone.sum() > 0


def test_policy_kernel_matches_reference_legal_lists(backend):
    """Every legal list position is reachable and maps to the reference's action (incl. Rock(15,15)'s doubled rock)."""
    for board, k in [(7, 8), (15, 15)]:
        N = 4096
        env = gp.make("Rock-v0", board_size=board, num_rocks=k, batch_size=N, device=backend, seed=SEED)
        rs = np.random.RandomState(1)
        x, y, status = rs.randint(0, board, N), rs.randint(0, board, N), rs.randint(-1, 2, (N, k))
        a = env.sample_legal_actions(env.pack(x, y, status), step_ctr=3).cpu().numpy()
        w = philox.draw_slots(SEED, np.arange(N), 3, philox.DOMAIN_POLICY, 1)[:, 0]
        for i in range(0, N, 7):
            if board == 15 and (x[i], y[i]) == (12, 2):
                continue
            legal = C.rock_legal(board, k, x[i], y[i], status[i])
            assert legal == O.rock_generate_legal(O.RockCfg(board, k), int(x[i]), int(y[i]), status[i].tolist())
            assert a[i] == legal[(int(w[i]) * len(legal)) >> 32], i
        mask = env._generate_legal(env.pack(x, y, status)).cpu().numpy()
        assert mask[np.arange(N), a].all()


@pytest.mark.parametrize("name,kw", [("Rock-v0", dict(board_size=11, num_rocks=11)), ("Rock-v0", dict(board_size=15, num_rocks=15)),
                                     ("StochasticRock-v0", {}), ("Tag-v0", {}), ("Tag-v0", dict(num_opponents=3)),
                                     ("Tiger-v0", {}), ("Network-v0", {}), ("Battleship-v0", dict(board_size=(10, 10)))])
@pytest.mark.parametrize("goff", [0, 3])
def test_fused_rollout_equals_policy_plus_step_launches(backend, name, kw, goff):
    """goff = 0: the four-envs-per-thread vector kernels; goff = 3: the scalar kernels (+ odd views)."""
    N, T = 1001, 40
    env = gp.make(name, batch_size=N, device=backend, seed=77, global_offset=goff, **kw)
    s0, _ = env.init_states(N, step_ctr=1)
    final, ret, steps, flags = env.rollout(s0, max_steps=T, step_ctr=10)
    s = s0
    r = torch.zeros(N, dtype=torch.float64)     # accumulated on the CPU: torch's CUDA `x / 10` multiplies by a reciprocal
    st = torch.zeros(N, dtype=torch.int32, device=backend)
    fl = torch.zeros(N, dtype=torch.int32, device=backend)
    disc = 1.0
    for t in range(T):
        a = env.sample_legal_actions(s, step_ctr=10 + t)
        ns, ob, rw, f = env.simulate(s, a, step_ctr=10 + t)
        act = (f & _lib.FLAG_STEPPED_DONE) == 0
        rw64 = torch.round(rw.double().cpu() * 10) / 10 if name == "Network-v0" else rw.double().cpu()
        r = torch.where(act.cpu(), r + rw64 * disc, r)
        st += act.int()
        fl |= torch.where(act, f, torch.zeros_like(f))
        s, disc = ns, disc * env._discount
    assert torch.equal(final, s), (final != s.reshape(final.shape)).nonzero()[:8].tolist()
    assert torch.equal(steps, st) and torch.equal(flags, fl)
    bad = (ret.cpu() != r).nonzero()[:4, 0].tolist()
    assert not bad, [(i, float(ret[i]).hex(), float(r[i]).hex()) for i in bad]
    # an env that is already terminal takes no step
    f2, r2, s2, fl2 = env.rollout(final, max_steps=5, step_ctr=100)
    fin = (flags & 1) != 0
    assert torch.equal(f2[fin], final[fin]) and not r2[fin].any() and not s2[fin].any() and (fl2[fin] == 1).all()
    # max_steps = 0 is the identity
    f3, r3, s3, fl3 = env.rollout(s0, max_steps=0, step_ctr=10)
    assert torch.equal(f3, s0) and not r3.any() and not s3.any()


@pytest.mark.parametrize("name,kw", [("Rock-v0", dict(board_size=11, num_rocks=11)), ("Tag-v0", {}), ("Tiger-v0", {}),
                                     ("Network-v0", {}), ("Battleship-v0", dict(board_size=(10, 10)))])
def test_rollout_with_first_action_is_simulate_then_rollout(backend, name, kw):
    """Q(s, a) samples: ``rollout(first_action=a)`` == ``simulate(s, a)`` at counter c, then a policy rollout from c + 1,
    with the first reward undiscounted and the rest discounted once more -- and it matches the C oracle."""
    N, T = 1001, 25
    env = gp.make(name, batch_size=N, device=backend, seed=SEED, **kw)
    s0, _ = env.init_states(N, step_ctr=1)
    rs = np.random.RandomState(5)
    a0 = torch.as_tensor(rs.randint(0, env.action_space.n, N), device=backend).int()
    final, ret, steps, flags = env.rollout(s0, max_steps=T, step_ctr=10, first_action=a0)
    s1, _, r1, f1 = env.simulate(s0, a0, step_ctr=10)
    f2, ret2, steps2, fl2 = env.rollout(s1, max_steps=T - 1, step_ctr=11)
    assert torch.equal(final, f2) and torch.equal(steps, steps2 + 1) and torch.equal(flags, f1 | fl2)
    if name == "Rock-v0":
        # same value through the oracle (the Python-side recombination r1 + gamma * ret2 rounds differently)
        x, y, st, _ = (v.cpu().numpy() for v in env.unpack(s0))
        ex, ey, est, eret, esteps, edone, err = C.rock_rollout(11, 11, False, 0.8, x, y, st, SEED, 0, 10, T, env._discount,
                                                               first_action=a0.cpu().numpy())
        assert np.array_equal(ret.cpu().numpy(), eret) and np.array_equal(steps.cpu().numpy(), esteps)
    elif name == "Tag-v0":
        ag, op, nop, _ = (v.cpu().numpy() for v in env.unpack(s0))
        _, _, _, eret, esteps, _ = C.tag_rollout(1, 0.8, ag, op, nop, SEED, 0, 10, T, env._discount, first_action=a0.cpu().numpy())
        assert np.array_equal(ret.cpu().numpy(), eret) and np.array_equal(steps.cpu().numpy(), esteps)
    else:
        r1d = torch.round(r1.double().cpu() * 10) / 10 if name == "Network-v0" else r1.double().cpu()
        approx = r1d + env._discount * ret2.cpu() * ((f1.cpu() & 1) == 0)
        assert torch.allclose(ret.cpu(), approx, rtol=1e-12, atol=1e-9)
