"""BASELINE.json's configurations at their FULL batch sizes, element-wise against the C oracle
(oracle/pomdp_oracle.c: the plain-C restatement of the reference, pinned to the reference's
fixtures by tests/test_oracle_c_golden.py) fed the same Philox words -- every env instance
of every batch is compared: next state, observation, reward, done.  Bar: bit-exact.

  RockSample(7,8)  B = 2^20      Tag-v0  B = 2^20      BattleShip 10x10  B = 2^18
  RockSample(11,11) B = 2^22 (the metric's config)
  RockSample(15,15): one 2^22 shard of the 2^25 batch at its global offset, plus the whole
                     2^25 batch on one device checked shard-against-whole.

The ``gpu`` run uses those sizes through the CUDA library; the CPU suite runs the same code
at 2^14 on the host build of the functors.
"""
import numpy as np
import pytest
import torch

import gym_pomdp_b200 as gp
from gym_pomdp_b200 import _lib
from oracle import c_oracle as C
from oracle import philox

from backends import backend  # noqa: F401

SEED = 0x5EED


def size(backend, log2):
    return 1 << (log2 if backend.startswith("cuda") else 14)


def dev_ints(gen, lo, hi, shape, dev):
    return torch.randint(lo, hi, shape, generator=gen, device=dev)


def gen_for(dev, salt):
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + salt)
    return g


def rock_case(backend, board, k, B, goff, stochastic=False, ctr=17):
    env = gp.make("StochasticRock-v0" if stochastic else "Rock-v0", board_size=board, num_rocks=k, batch_size=B,
                  device=backend, seed=SEED, global_offset=goff)
    g = gen_for(backend, board)
    x, y = dev_ints(g, 0, board, (B,), backend), dev_ints(g, 0, board, (B,), backend)
    status = dev_ints(g, -1, 2, (B, k), backend)
    action = dev_ints(g, 0, 5 + k, (B,), backend).int()
    state = env.pack(x, y, status)
    ns, ob, rw, fl = env.simulate(state, action, step_ctr=ctr)
    x2, y2, st2, done = (v.cpu().numpy() for v in env.unpack(ns))
    draws = C.fill_draws(SEED, goff, B, ctr, philox.DOMAIN_STEP, 2)
    ex, ey, est, eob, erw, edone, err = C.rock_step(board, k, stochastic, 0.8, x.cpu().numpy(), y.cpu().numpy(),
                                                   status.cpu().numpy(), action.cpu().numpy(), draws)
    fl = fl.cpu().numpy()
    assert np.array_equal(x2, ex) and np.array_equal(y2, ey)
    assert np.array_equal(st2, est.astype(np.int32))
    assert np.array_equal(ob.cpu().numpy(), eob)
    assert np.array_equal(rw.cpu().numpy(), erw.astype(np.float32))
    assert np.array_equal(done, edone) and np.array_equal((fl & 1).astype(bool), edone)
    assert np.array_equal((fl & _lib.FLAG_BAD_STATE) != 0, err != 0)       # the dangling cell of Rock(15,15) / Rock(7,7)
    assert not (fl & (_lib.FLAG_BAD_ACTION | _lib.FLAG_STEPPED_DONE)).any()
    # reset of the same shard
    st0, ob0 = env.init_states(B, step_ctr=ctr + 1)
    rx, ry, rst, rdone = (v.cpu().numpy() for v in env.unpack(st0))
    qx, qy, qst, qob = C.rock_reset(board, k, C.fill_draws(SEED, goff, B, ctr + 1, philox.DOMAIN_RESET, 1))
    assert np.array_equal(rx, qx) and np.array_equal(ry, qy) and np.array_equal(rst, qst.astype(np.int32))
    assert not ob0.any() and not rdone.any()
    return env, state, action, (ns, ob, rw, torch.as_tensor(fl))


def test_rock_7_8_batch_2p20(backend):
    rock_case(backend, 7, 8, size(backend, 20), 0)


def test_rock_11_11_batch_2p22(backend):
    rock_case(backend, 11, 11, size(backend, 22), 0)


def test_stochastic_rock_11_11_batch_2p20(backend):
    rock_case(backend, 11, 11, size(backend, 20), 0, stochastic=True)


def test_rock_15_15_one_shard_of_2p25(backend):
    B = size(backend, 22)
    rock_case(backend, 15, 15, B, 5 * B)           # rank 5 of 8


def test_rock_15_15_whole_2p25_equals_its_eight_shards(backend):
    """The full config-5 batch on one device vs the same batch cut into the eight index shards
    the 8-GPU run uses (global_offset = rank * B/8): identical words everywhere."""
    B = size(backend, 25)
    env = gp.make("Rock-v0", board_size=15, num_rocks=15, batch_size=B, device=backend, seed=SEED)
    g = gen_for(backend, 99)
    state = env.pack(dev_ints(g, 0, 15, (B,), backend), dev_ints(g, 0, 15, (B,), backend),
                     dev_ints(g, -1, 2, (B, 15), backend))
    action = dev_ints(g, 0, 20, (B,), backend).int()
    whole = env.simulate(state, action, step_ctr=3)
    per = B // 8
    for r in range(8):
        e = gp.make("Rock-v0", board_size=15, num_rocks=15, batch_size=per, device=backend, seed=SEED,
                    global_offset=r * per)
        part = e.simulate(state[r * per:(r + 1) * per], action[r * per:(r + 1) * per], step_ctr=3)
        assert all(torch.equal(p, w[r * per:(r + 1) * per]) for p, w in zip(part, whole)), r
    ns, ob, rw, fl = whole
    # value domains (rock.py:123-194): obs in {0,1,2}; reward in {0, +10, -10, -100}; done <=> reward -100 or east exit
    assert int(ob.min()) >= 0 and int(ob.max()) <= 2
    assert set(torch.unique(rw).tolist()) <= {0.0, 10.0, -10.0, -100.0}
    done = (fl & 1).bool()
    x, _, _, d2 = env.unpack(ns)
    assert torch.equal(done, d2)
    assert torch.equal(done, (rw == -100) | ((rw == 10) & (action == 1) & (x == 14)))


def test_tag_batch_2p20(backend):
    B = size(backend, 20)
    for n_opp in (1, 2):
        env = gp.make("Tag-v0", num_opponents=n_opp, batch_size=B, device=backend, seed=SEED)
        g = gen_for(backend, 30 + n_opp)
        agent, opp = dev_ints(g, 0, 29, (B,), backend), dev_ints(g, 0, 29, (B, n_opp), backend)
        action = dev_ints(g, 0, 5, (B,), backend).int()
        ns, ob, rw, fl = env.simulate(env.pack(agent, opp), action, step_ctr=23)
        a2, o2, nop2, done = (v.cpu().numpy() for v in env.unpack(ns))
        draws = C.fill_draws(SEED, 0, B, 23, philox.DOMAIN_STEP, n_opp)
        ea, eo, enop, eob, erw, edone = C.tag_step(n_opp, 0.8, agent.cpu().numpy(), opp.cpu().numpy(),
                                                   np.full(B, n_opp, np.int32), action.cpu().numpy(), draws)
        assert np.array_equal(a2, ea) and np.array_equal(o2, eo) and np.array_equal(nop2, enop)
        assert np.array_equal(ob.cpu().numpy(), eob) and np.array_equal(rw.cpu().numpy(), erw.astype(np.float32))
        assert np.array_equal(done, edone) and np.array_equal(fl.cpu().numpy(), edone.astype(np.int32))
        st0, ob0 = env.init_states(B, step_ctr=24)
        ra, ro, rn, _ = (v.cpu().numpy() for v in env.unpack(st0))
        qa, qo, qn, qob = C.tag_reset(n_opp, C.fill_draws(SEED, 0, B, 24, philox.DOMAIN_RESET, (1 + n_opp + 2) // 3))
        assert np.array_equal(ra, qa) and np.array_equal(ro, qo) and np.array_equal(rn, qn)
        assert np.array_equal(ob0.cpu().numpy(), qob)


def test_tiger_and_network_batch_2p20(backend):
    B = size(backend, 20)
    env = gp.make("Tiger-v0", batch_size=B, device=backend, seed=SEED)
    g = gen_for(backend, 40)
    s0, action = dev_ints(g, 0, 2, (B,), backend), dev_ints(g, 0, 3, (B,), backend).int()
    ns, ob, rw, fl = env.simulate(env.pack(s0), action, step_ctr=5)
    s2, done = (v.cpu().numpy() for v in env.unpack(ns))
    es, eob, erw, edone = C.tiger_step(0.85, s0.cpu().numpy(), action.cpu().numpy(),
                                       C.fill_draws(SEED, 0, B, 5, philox.DOMAIN_STEP, 1))
    assert np.array_equal(s2, es) and np.array_equal(ob.cpu().numpy(), eob) and np.array_equal(done, edone)
    assert np.array_equal(rw.cpu().numpy(), erw.astype(np.float32))
    st0, ob0 = env.init_states(B, step_ctr=6)
    qs, qob = C.tiger_reset(C.fill_draws(SEED, 0, B, 6, philox.DOMAIN_RESET, 1))
    assert np.array_equal(env.unpack(st0)[0].cpu().numpy(), qs) and np.array_equal(ob0.cpu().numpy(), qob)

    for n, ptype in [(10, 3), (16, 3), (12, 0), (28, 3), (30, 0)]:
        env = gp.make("Network-v0", n_machines=n, problem_type=ptype, batch_size=B, device=backend, seed=SEED)
        s0 = dev_ints(g, 0, 1 << n, (B,), backend)
        action = dev_ints(g, 0, 2 * n + 1, (B,), backend).int()
        ns, ob, rw, fl = env.simulate(s0.int(), action, step_ctr=7)
        bits = ((s0.cpu().numpy()[:, None] >> np.arange(n)) & 1).astype(np.int8)
        em, eob, erw = C.network_step(n, ptype, bits, action.cpu().numpy(), C.network_draws(SEED, 0, B, 7, n))
        assert np.array_equal(ns.cpu().numpy(), (em.astype(np.int64) << np.arange(n)).sum(1).astype(np.int32))
        assert np.array_equal(ob.cpu().numpy(), eob)
        assert np.array_equal(rw.cpu().numpy(), erw.astype(np.float32))       # float32(the reference's double)
        assert not fl.any()


def test_battleship_10x10_batch_2p18(backend):
    """Warp-per-board placement scan and the shot kernel at B = 2^18."""
    B = size(backend, 18)
    env = gp.make("Battleship-v0", board_size=(10, 10), batch_size=B, device=backend, seed=SEED)
    st, ob0 = env.init_states(B, step_ctr=2)
    occ, vis, rem, done = (v.cpu().numpy() for v in env.unpack(st))
    eocc, erem, err = C.battleship_reset_scan(10, 10, 3, C.fill_env_draws(SEED, 0, B, 2, philox.DOMAIN_SHIP, 2))
    assert np.array_equal(occ.reshape(B, 10, 10), eocc) and np.array_equal(rem, erem) and not err.any()
    assert not vis.any() and not done.any() and not ob0.any() and not env.reset_flags.any()
    # synthetic visited pattern ~ Bernoulli(0.3), total_remaining recomputed, then four shots
    g = gen_for(backend, 50)
    visited = torch.rand((B, 10, 10), generator=g, device=backend) < 0.3
    occ_t = torch.as_tensor(eocc, device=backend)
    remaining = (occ_t & ~visited).reshape(B, -1).sum(1)
    keep = remaining > 0                                             # battleship.py:95: total_remaining > 0
    visited[~keep] = False
    remaining = (occ_t & ~visited).reshape(B, -1).sum(1)
    state = env.pack(occ_t, visited, total_remaining=remaining)
    vis_np, rem_np = visited.cpu().numpy(), remaining.cpu().numpy().astype(np.int32)
    alive = np.ones(B, bool)
    for s in range(4):
        action = dev_ints(g, 0, 100, (B,), backend).int()
        ns, ob, rw, fl = env.simulate(state, action, step_ctr=10 + s)
        vis_np, rem_np, eob, erw, edone = C.battleship_step(10, 10, eocc, vis_np, rem_np, action.cpu().numpy())
        _, v2, r2, d2 = (v.cpu().numpy() for v in env.unpack(ns))
        assert np.array_equal(ob.cpu().numpy()[alive], eob[alive])
        assert np.array_equal(rw.cpu().numpy()[alive], erw[alive].astype(np.float32))
        assert np.array_equal(d2[alive], edone[alive]) and np.array_equal(r2[alive], rem_np[alive])
        assert np.array_equal(v2.reshape(B, 10, 10)[alive], vis_np[alive])
        fin = ~alive
        assert ((fl.cpu().numpy()[fin] & _lib.FLAG_STEPPED_DONE) != 0).all()
        alive &= ~edone
        # the oracle keeps mutating finished boards; re-sync them from the device for the next shot
        vis_np[~alive] = v2.reshape(B, 10, 10)[~alive]
        rem_np[~alive] = r2[~alive]
        state = ns


def test_belief_histogram_2p22(backend):
    B = size(backend, 22)
    env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=B, device=backend, seed=SEED)
    st, _ = env.init_states(B, step_ctr=1)
    h = env.belief_histogram(st).cpu().numpy()
    x, y, status, _ = (v.cpu().numpy() for v in env.unpack(st))
    assert np.array_equal(h[:11], (status == 1).sum(0)) and h[11 + (0 | 5 << 4)] == B and h[11:].sum() == B


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["rock15", "rock11", "network"])
def test_belief_histogram_carry_save_flavour_at_2p25(which):
    """Batches large enough for the kernel's carry-save (Harley-Seal) counting of the bit bins -- every thread makes at least
    eight trips -- with a ragged end (n % 4 != 0: the scalar tail follows the vector loop) and random states, against
    torch's own reductions over the unpacked states."""
    dev = "cuda:0"
    n = (1 << 25) + 3
    g = gen_for(dev, 91)
    if which == "network":
        env = gp.make("Network-v0", n_machines=30, problem_type=0, batch_size=n, device=dev, seed=SEED)
        st = torch.randint(0, 1 << 30, (n,), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
        exp = torch.stack([((st >> m) & 1).sum() for m in range(30)])
    else:
        b, k = (15, 15) if which == "rock15" else (11, 11)
        env = gp.make("Rock-v0", board_size=b, num_rocks=k, batch_size=n, device=dev, seed=SEED)
        x = torch.randint(0, b, (n,), generator=g, device=dev)
        y = torch.randint(0, b, (n,), generator=g, device=dev)
        status = torch.randint(-1, 2, (n, k), generator=g, device=dev, dtype=torch.int8)
        st = env.pack(x, y, status)
        exp = torch.cat([(status == 1).sum(0), torch.bincount(x | (y << 4), minlength=256)])
        del x, y, status
    h = env.belief_histogram(st)
    assert torch.equal(h, exp.to(h.dtype)), which
    # and a view that starts 16 bytes in (still vector-aligned) but ends ragged
    off = 4 * env.state_words
    h2 = env.belief_histogram(st[4:]) if env.state_words == 1 else env.belief_histogram(st[2:])
    assert int(h2.sum()) < int(h.sum()) and off > 0


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["rock15_2p25", "rock11_2p26", "tag_2p20", "network_2p22", "tiger_2p22"])
def test_step_with_histogram_epilogue_at_baseline_sizes(which):
    """pomdp_E_step_hist at the BASELINE batch sizes (the whole Rock(15,15) 2^25 batch of config 5 included, ragged by 3):
    every output array equal to the plain step's, the counts equal to torch's own reductions over the unpacked next states.
    Rock(11,11) runs at 2^26: a thread then counts 4 x 111 particles, so the byte counters are flushed INSIDE the loop."""
    dev = "cuda:0"
    g = gen_for(dev, 17)
    if which.startswith("rock"):
        b, k, lg = (15, 15, 25) if which == "rock15_2p25" else (11, 11, 26)
        n = (1 << lg) + 3
        env = gp.make("Rock-v0", board_size=b, num_rocks=k, batch_size=n, device=dev, seed=SEED)
        st = env.pack(torch.randint(0, b, (n,), generator=g, device=dev), torch.randint(0, b, (n,), generator=g, device=dev),
                      torch.randint(-1, 2, (n, k), generator=g, device=dev, dtype=torch.int8))
        act = torch.randint(0, 5 + k, (n,), generator=g, device=dev, dtype=torch.int32)
    elif which.startswith("tag"):
        n = (1 << 20) + 3
        env = gp.make("Tag-v0", batch_size=n, device=dev, seed=SEED)
        st = env.pack(torch.randint(0, 29, (n,), generator=g, device=dev), torch.randint(0, 29, (n, 1), generator=g, device=dev))
        act = torch.randint(0, 5, (n,), generator=g, device=dev, dtype=torch.int32)
    elif which.startswith("network"):
        n = (1 << 22) + 3
        env = gp.make("Network-v0", batch_size=n, device=dev, seed=SEED)
        st = torch.randint(0, 1024, (n,), generator=g, device=dev, dtype=torch.int32)
        act = torch.randint(0, 21, (n,), generator=g, device=dev, dtype=torch.int32)
    else:
        n = (1 << 22) + 3
        env = gp.make("Tiger-v0", batch_size=n, device=dev, seed=SEED)
        st = env.pack(torch.randint(0, 2, (n,), generator=g, device=dev))
        act = torch.randint(0, 3, (n,), generator=g, device=dev, dtype=torch.int32)
    ref = env.simulate(st, act, step_ctr=3)
    got = env.simulate_hist(st, act, step_ctr=3)
    assert all(torch.equal(a, b) for a, b in zip(got[:4], ref)), which
    ns = ref[0]
    if which.startswith("rock"):
        x, y, status, _ = env.unpack(ns)
        exp = torch.cat([(status == 1).sum(0), torch.bincount((x | (y << 4)).long(), minlength=256)])
    elif which.startswith("tag"):
        ag, op, _, _ = env.unpack(ns)
        exp = torch.cat([torch.bincount(ag.long(), minlength=29), torch.bincount(op[:, 0].long(), minlength=29)])
    elif which.startswith("network"):
        exp = torch.stack([((ns >> m) & 1).sum() for m in range(10)])
    else:
        exp = torch.bincount((ns & 1).long(), minlength=2)
    assert torch.equal(got[4], exp.to(got[4].dtype)), which
    assert torch.equal(env.simulate_hist(st, act, step_ctr=3)[4], got[4])          # the scratch was left clean


@pytest.mark.gpu
def test_batch_beyond_2p31_envs():
    """Maximum sizes: 2^31 + 4 Tiger instances in ONE launch (52 GB of arrays): element indices and Philox counters are
    64-bit end to end.  Windows at the start, around 2^31 and at the ragged end are checked against the C oracle."""
    dev = "cuda:0"
    free, _ = torch.cuda.mem_get_info()
    n = (1 << 31) + 4
    if free < 60 * (1 << 30):
        pytest.skip("needs 60 GB of free device memory")
    env = gp.make("Tiger-v0", batch_size=n, device=dev, seed=SEED)
    g = gen_for(dev, 77)
    state = torch.randint(0, 2, (n,), generator=g, device=dev, dtype=torch.int32)
    action = torch.randint(0, 3, (n,), generator=g, device=dev, dtype=torch.int32)
    ns, ob, rw, fl = env.simulate(state, action, step_ctr=3)
    torch.cuda.synchronize()
    for lo, hi in [(0, 4096), ((1 << 31) - 4096, (1 << 31) + 4), (n - 3, n)]:
        es, eob, erw, edone = C.tiger_step(0.85, state[lo:hi].cpu().numpy(), action[lo:hi].cpu().numpy(),
                                           C.fill_draws(SEED, lo, hi - lo, 3, philox.DOMAIN_STEP, 1))
        s2, done = (v.cpu().numpy() for v in env.unpack(ns[lo:hi]))
        assert np.array_equal(s2, es) and np.array_equal(ob[lo:hi].cpu().numpy(), eob) and np.array_equal(done, edone)
        assert np.array_equal(rw[lo:hi].cpu().numpy(), erw.astype(np.float32))
    # every element was written: obs in {0, 1, 2}, flags in {0, 1}
    assert int(ob.max()) <= 2 and int(ob.min()) >= 0 and int(fl.max()) <= 1 and int(fl.min()) >= 0
    del ns, ob, rw, fl, state, action
    torch.cuda.empty_cache()
