"""Pins oracle/pomdp_oracle.c (the plain-C restatement) to the fixtures oracle/gen_golden.py
recorded from the unmodified reference.  CPU only; whole-array comparisons."""
import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import philox

ROCKS = ["7_8", "11_11", "15_15", "7_7", "4_3", "stoch_7_8", "stoch_11_11"]


def test_philox_kat_and_fill():
    vecs = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in vecs:   # Random123 known-answer vectors
        assert C.philox(ctr, key).tolist() == list(exp)
    for (seed, off, step, dom, slots) in [(0x5EED, 0, 11, 0, 2), (2 ** 63 + 5, 2 ** 33 + 12, 7, 1, 11), (1, 3, 0, 1, 96)]:
        a = C.fill_draws(seed, off, 257, step, dom, slots)
        b = philox.draw_slots(seed, np.arange(257, dtype=np.uint64) + np.uint64(off), step, dom, slots)
        assert np.array_equal(a, b)
    # env-keyed draws (BattleShip's fixed-time placement): one block = four consecutive slots of ONE env
    for (seed, off, step, slots) in [(0x5EED, 0, 2, 2), (2 ** 63 + 5, 2 ** 33 + 12, 7, 8), (1, 3, 0, 5)]:
        a = C.fill_env_draws(seed, off, 257, step, philox.DOMAIN_SHIP, slots)
        b = philox.draw_env_slots(seed, np.arange(257, dtype=np.uint64) + np.uint64(off), step, philox.DOMAIN_SHIP, slots)
        assert np.array_equal(a, b)
        blk = C.philox((int(off) & 0xFFFFFFFF, int(off) >> 32, step, philox.DOMAIN_SHIP << 24), (seed & 0xFFFFFFFF, seed >> 32))
        assert a[0, :min(4, slots)].tolist() == blk.tolist()[:min(4, slots)]


@pytest.mark.parametrize("tag", ROCKS)
def test_rock(golden, tag):
    g = golden("rock_" + tag)
    n, k, stoch = int(g["n"]), int(g["k"]), bool(g["stochastic"])
    grid, pos, start = C.rock_grid(n, k)
    assert np.array_equal(grid, g["grid"]) and np.array_equal(pos, g["rock_pos"]) and np.array_equal(start, g["start"])
    assert [C.rock_efficiency(d) for d in range(len(g["eff"]))] == g["eff"].tolist()   # bit-equal doubles
    x2, y2, st2, ob, rw, done, err = C.rock_step(n, k, stoch, 0.8, g["x"], g["y"], g["status"], g["action"], g["draws"])
    ok = ~g["raised"]
    assert np.array_equal(err[g["raised"]], np.full(int(g["raised"].sum()), 8)) and not err[ok].any()
    assert np.array_equal(x2[ok], g["x2"][ok]) and np.array_equal(y2[ok], g["y2"][ok])
    assert np.array_equal(st2[ok], g["status2"][ok])
    assert np.array_equal(ob[ok], g["obs"][ok]) and np.array_equal(rw[ok], g["reward"][ok])
    assert np.array_equal(done[ok], g["done"][ok])
    rx, ry, rst, rob = C.rock_reset(n, k, g["reset_draws"])
    assert np.array_equal(np.stack([rx, ry], 1), g["reset_xy"]) and np.array_equal(rst, g["reset_status"])
    assert np.array_equal(rob, g["reset_obs"])


@pytest.mark.parametrize("tag", ["1opp", "2opp"])
def test_tag(golden, tag):
    g = golden("tag_" + tag)
    n_opp = int(g["n_opp"])
    adm = np.array([[C.tag_admissible(a, o) for o in range(29)] for a in range(29)])
    assert np.array_equal(adm, g["admissible"])
    agent2, opp2, nop2, ob, rw, done = C.tag_step(n_opp, float(g["move_prob"]), g["agent"], g["opp"], g["num_opp"],
                                                  g["action"], g["draws"])
    assert np.array_equal(agent2, g["agent2"]) and np.array_equal(opp2, g["opp2"]) and np.array_equal(nop2, g["num_opp2"])
    assert np.array_equal(ob, g["obs"]) and np.array_equal(rw, g["reward"]) and np.array_equal(done, g["done"])
    ra, ro, rn, rob = C.tag_reset(n_opp, g["reset_draws"])
    assert np.array_equal(ra, g["reset_agent"]) and np.array_equal(ro, g["reset_opp"]) and np.array_equal(rob, g["reset_obs"])
    assert (rn == n_opp).all()


@pytest.mark.parametrize("tag", ["10x10", "5x5"])
def test_battleship(golden, tag):
    g = golden("battleship_" + tag)
    xs, ys, max_len = int(g["x_size"]), int(g["y_size"]), int(g["max_len"])
    B = len(g["occupied"])
    occ, ships, attempts, rem = C.battleship_reset_rejection(xs, ys, max_len, g["reset_draws"])
    assert np.array_equal(occ, g["occupied"]) and np.array_equal(attempts, g["attempts"]) and (rem == 5).all()
    assert np.array_equal(ships, g["ships"])
    v1, c1 = C.battleship_valid(xs, ys, np.zeros((xs, ys), np.uint8), max_len)
    assert np.array_equal(v1, g["valid_first"]) and c1 == int(g["valid_first"].sum())
    for b in range(B):   # accepted set for the 2nd ship given the 1st
        first = np.zeros((xs, ys), np.uint8)
        x, y, d, ln = (int(v) for v in g["ships"][b, 0])
        dx, dy = [(0, 1), (1, 0), (0, -1), (-1, 0)][d]
        for i in range(ln):
            first[x + i * dx, y + i * dy] = 1
        v2, _ = C.battleship_valid(xs, ys, first, max_len - 1)
        assert np.array_equal(v2, g["valid_second"][b]), b
    vis, rem = g["visited_in"].copy(), g["remaining"][:, 0].copy()
    alive = np.ones(B, bool)
    for s in range(g["action"].shape[1]):
        a = g["action"][:, s]
        alive &= a >= 0
        vis, rem, ob, rw, done = C.battleship_step(xs, ys, g["occupied"], vis, rem, np.where(a >= 0, a, 0))
        assert np.array_equal(ob[alive], g["obs"][alive, s]) and np.array_equal(rw[alive], g["reward"][alive, s])
        assert np.array_equal(done[alive], g["done"][alive, s]) and np.array_equal(rem[alive], g["remaining"][alive, s + 1])
    # the fixed-time scan reset draws from exactly the accepted set: kth valid candidate in order
    w = C.fill_draws(7, 0, 64, 0, 1, 2)
    occ_s, rem_s, err_s = C.battleship_reset_scan(xs, ys, max_len, w)
    assert (occ_s.reshape(64, -1).sum(1) == 5).all() and (rem_s == 5).all() and not err_s.any()
    valid_idx = np.nonzero(g["valid_first"])[0]
    for b in range(64):
        c = int(valid_idx[(int(w[b, 0]) * len(valid_idx)) >> 32])
        x, y, d = (c >> 2) % xs, (c >> 2) // xs, c & 3
        dx, dy = [(0, 1), (1, 0), (0, -1), (-1, 0)][d]
        assert all(occ_s[b, x + i * dx, y + i * dy] for i in range(max_len))


def test_tiger(golden):
    g = golden("tiger")
    s2, ob, rw, done = C.tiger_step(0.85, g["state"], g["action"], g["draws"])
    assert np.array_equal(s2, g["state2"]) and np.array_equal(ob, g["obs"]) and np.array_equal(rw, g["reward"])
    assert np.array_equal(done, g["done"])
    rs, rob = C.tiger_reset(g["reset_draws"])
    assert np.array_equal(rs, g["reset_state"]) and np.array_equal(rob, g["reset_obs"])


@pytest.mark.parametrize("tag", ["3legs10", "3legs7", "ring10", "3legs19"])
def test_network(golden, tag):
    g = golden("network_" + tag)
    n, pt = int(g["n"]), int(g["problem_type"])
    assert np.array_equal(C.network_neighbours(n, pt), g["neighbours"])
    bits = ((g["state"][:, None] >> np.arange(n)) & 1).astype(np.int8)
    m2, ob, rw = C.network_step(n, pt, bits, g["action"], g["draws"], float(g["p"]), float(g["q"]), float(g["p_ob"]))
    assert np.array_equal((m2.astype(np.int64) << np.arange(n)).sum(1), g["state2"])
    assert np.array_equal(ob, g["obs"])
    assert np.array_equal(rw, g["reward"])     # the reference's double, bit for bit
