"""Pins oracle/pomdp_oracle.py (the pure-Python restatement) to the fixtures that
oracle/gen_golden.py recorded from the unmodified reference.  CPU only."""
import numpy as np
import pytest

from oracle import philox
from oracle import pomdp_oracle as O

ROCKS = ["7_8", "11_11", "15_15", "7_7", "4_3", "stoch_7_8", "stoch_11_11"]


def test_philox_kat():
    assert philox.kat()


def test_coord_kats(golden):
    g = golden("coord")
    # the reference's own six asserts (coord.py:121-126)
    assert ((g["kat_a"] + g["kat_b"]) == g["kat_sum"]).all()
    assert [tuple(m) for m in g["moves"]] == list(O.MOVES)
    assert [O.grid_opposite(m) for m in range(4)] == g["opposite"].tolist()
    for (xs, ys) in [(7, 7), (11, 11), (15, 15), (10, 10), (5, 5), (10, 5)]:
        coords = g[f"grid_{xs}x{ys}_coord"]
        for idx, (x, y) in enumerate(coords):
            assert O.grid_get_coord(xs, idx) == (x, y)
            assert O.grid_get_index(xs, x, y) == idx
        for (x, y), inside in zip(g[f"grid_{xs}x{ys}_probe"], g[f"grid_{xs}x{ys}_inside"]):
            assert O.grid_is_inside(xs, ys, x, y) == inside
    for idx, (x, y) in enumerate(g["tag_coord"]):
        assert O.tag_get_coord(idx) == (x, y)
        assert O.tag_get_index(x, y) == g["tag_index"][idx] == idx
    for (x, y), inside in zip(g["tag_probe"], g["tag_inside"]):
        assert O.tag_is_inside(x, y) == inside
    pts = g["dist_pts"]
    for i, a in enumerate(pts):
        for j, b in enumerate(pts):
            assert O.l1_distance(*a, *b) == g["dist_l1"][i, j]


@pytest.mark.parametrize("tag", ROCKS)
def test_rock_step(golden, tag):
    g = golden("rock_" + tag)
    cfg = O.RockCfg(int(g["n"]), int(g["k"]), bool(g["stochastic"]))
    assert np.array_equal(np.array(cfg.grid, np.int8), g["grid"])
    assert [cfg.efficiency(d) for d in range(cfg.max_dist + 1)] == g["eff"].tolist()  # bit-equal doubles
    assert cfg.n_actions == int(g["n_actions"])
    n_raised = 0
    for i in range(len(g["x"])):
        used = []

        def draw(slot, i=i):
            used.append(slot)
            return int(g["draws"][i, slot])
        x, y, st, ob, rw, done, err = O.rock_step(cfg, int(g["x"][i]), int(g["y"][i]), g["status"][i].tolist(),
                                                  int(g["action"][i]), draw)
        if g["raised"][i]:
            n_raised += 1
            assert err & O.ROCK_ERR_DANGLING  # reference: IndexError (rock.py:162)
            continue
        assert err == 0
        assert (x, y) == (g["x2"][i], g["y2"][i]), i
        assert st == g["status2"][i].tolist(), i
        assert (ob, float(rw), done) == (g["obs"][i], g["reward"][i], bool(g["done"][i])), i
        assert len(used) == g["ndraw"][i], i  # same number of RNG calls as the reference
        for o in range(3):
            assert O.rock_compute_prob(cfg, int(g["action"][i]), x, y, st, o) == g["prob"][i, o]
        legal = [a for a in g["legal"][i].tolist() if a >= 0]
        if legal:
            assert O.rock_generate_legal(cfg, x, y, st) == legal
    assert n_raised == int(g["raised"].sum())


@pytest.mark.parametrize("tag", ROCKS)
def test_rock_reset(golden, tag):
    g = golden("rock_" + tag)
    cfg = O.RockCfg(int(g["n"]), int(g["k"]), bool(g["stochastic"]))
    for i in range(len(g["reset_draws"])):
        x, y, st, ob = O.rock_reset(cfg, lambda s, i=i: int(g["reset_draws"][i, s]))
        assert (x, y) == tuple(g["reset_xy"][i]) == tuple(g["start"])
        assert st == g["reset_status"][i].tolist()
        assert ob == g["reset_obs"][i] == 0
    assert g["reset_status"][0, 0] == 0  # the u == 0.5 corner really is in the fixture


@pytest.mark.parametrize("tag", ["1opp", "2opp"])
def test_tag(golden, tag):
    g = golden("tag_" + tag)
    for a in range(29):
        for o in range(29):
            exp = [m for m in g["admissible"][a, o].tolist() if m >= 0]
            assert O.tag_admissible(*O.tag_get_coord(a), *O.tag_get_coord(o)) == exp
    for i in range(len(g["agent"])):
        ax, ay = O.tag_get_coord(int(g["agent"][i]))
        opps = [O.tag_get_coord(int(o)) for o in g["opp"][i]]
        ax, ay, opps, nop, ob, rw, done = O.tag_step(ax, ay, opps, int(g["num_opp"][i]), int(g["action"][i]),
                                                     lambda s, i=i: int(g["draws"][i, s]), float(g["move_prob"]))
        assert O.tag_get_index(ax, ay) == g["agent2"][i], i
        assert [O.tag_get_index(*o) for o in opps] == g["opp2"][i].tolist(), i
        assert (nop, ob, rw, done) == (g["num_opp2"][i], g["obs"][i], g["reward"][i], bool(g["done"][i])), i
        for o in (0, int(ob), 28, 29):
            assert O.tag_compute_prob(ax, ay, opps, o) == g["prob"][i, o]
    n_opp = int(g["n_opp"])
    for i in range(len(g["reset_draws"])):
        ax, ay, opps, nop, ob = O.tag_reset(n_opp, lambda s, i=i: int(g["reset_draws"][i, s]))
        assert O.tag_get_index(ax, ay) == g["reset_agent"][i]
        assert [O.tag_get_index(*o) for o in opps] == g["reset_opp"][i].tolist()
        assert ob == g["reset_obs"][i]
    assert (g["reset_obs"] == 29).any()


@pytest.mark.parametrize("tag", ["10x10", "5x5"])
def test_battleship(golden, tag):
    g = golden("battleship_" + tag)
    xs, ys, max_len = int(g["x_size"]), int(g["y_size"]), int(g["max_len"])
    lengths = O.ship_lengths(max_len)
    empty = O.ShipBoard(xs, ys)
    valid1 = np.zeros(xs * ys * 4, bool)
    valid1[O.battleship_valid_placements(empty, lengths[0])] = True
    assert np.array_equal(valid1, g["valid_first"])
    for b in range(len(g["occupied"])):
        board, attempts = O.battleship_reset_rejection(xs, ys, max_len, lambda s, b=b: int(g["reset_draws"][b, s]))
        assert attempts == g["attempts"][b]
        assert np.array_equal(np.array(board.occupied), g["occupied"][b]), b
        assert board.total_remaining == sum(lengths)
        if b < 40:  # accepted set for the 2nd ship given the 1st (slow in pure Python)
            first = O.ShipBoard(xs, ys)
            x, y, d, ln = g["ships"][b, 0]
            O.ship_mark(first, int(x), int(y), int(d), int(ln))
            v2 = np.zeros(xs * ys * 4, bool)
            v2[O.battleship_valid_placements(first, lengths[1])] = True
            assert np.array_equal(v2, g["valid_second"][b]), b
        board.visited = g["visited_in"][b].tolist()
        board.total_remaining = int(g["remaining"][b, 0])
        for s in range(g["action"].shape[1]):
            a = int(g["action"][b, s])
            if a < 0:
                break
            ob, rw, done = O.battleship_step(board, a)
            assert (ob, rw, done) == (g["obs"][b, s], g["reward"][b, s], bool(g["done"][b, s])), (b, s)
            assert board.total_remaining == g["remaining"][b, s + 1]
            for o in range(2):
                assert O.battleship_compute_prob(board, a, o) == g["prob"][b, s, o]
            assert xs * ys - int(np.sum(board.visited)) == g["legal_count"][b, s]


def test_battleship_scan_matches_accepted_set(golden):
    """The fixed-time 'scan' reset picks from exactly the set the reference's rejection
    loop accepts (bit-exact valid masks above); here: it yields legal, complete boards."""
    for (xs, ys) in [(10, 10), (5, 5)]:
        for i in range(20):
            w = philox.draw_slots(7, [i], 0, philox.DOMAIN_RESET, 2)[0]
            board, ok = O.battleship_reset_scan(xs, ys, 3, lambda s: int(w[s]))
            assert ok and board.total_remaining == 5 and int(np.sum(board.occupied)) == 5


def test_tiger(golden):
    g = golden("tiger")
    for i in range(len(g["state"])):
        s2, ob, rw, done = O.tiger_step(int(g["state"][i]), int(g["action"][i]), lambda s, i=i: int(g["draws"][i, s]))
        assert (s2, ob, rw, done) == (g["state2"][i], g["obs"][i], g["reward"][i], bool(g["done"][i])), i
        for o in range(3):
            assert O.tiger_compute_prob(int(g["action"][i]), s2, o) == g["prob"][i, o]
    for i in range(len(g["reset_draws"])):
        s, ob = O.tiger_reset(lambda k, i=i: int(g["reset_draws"][i, k]))
        assert (s, ob) == (g["reset_state"][i], g["reset_obs"][i])


@pytest.mark.parametrize("tag", ["3legs10", "3legs7", "ring10", "3legs19"])
def test_network(golden, tag):
    g = golden("network_" + tag)
    n = int(g["n"])
    nb = O.network_neighbours(n, int(g["problem_type"]))
    exp_nb = [[j for j in row if j >= 0] for row in g["neighbours"].tolist()]
    assert nb == exp_nb
    step = 1 if len(g["state"]) < 6000 else 3  # keep the CPU suite quick
    for i in range(0, len(g["state"]), step):
        bits = [(int(g["state"][i]) >> m) & 1 for m in range(n)]
        st, ob, tenths, done = O.network_step(bits, int(g["action"][i]), lambda s, i=i: int(g["draws"][i, s]), nb,
                                              float(g["p"]), float(g["q"]), float(g["p_ob"]))
        assert sum(v << m for m, v in enumerate(st)) == g["state2"][i], i
        assert ob == g["obs"][i] and done is False
        assert tenths / 10.0 == g["reward"][i], i  # exact double equality with the reference's float
        for o in range(3):
            assert O.network_compute_prob(int(g["action"][i]), st, o) == g["prob"][i, o]
    st, ob = O.network_reset(n)
    assert sum(v << m for m, v in enumerate(st)) == int(g["reset_state"]) and ob == int(g["reset_obs"])
