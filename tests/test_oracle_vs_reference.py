"""Live cross-check of the oracle against the UNMODIFIED reference, when /root/reference is present
(the build container; skipped on the GPU box).  Random single steps, coupled through scripted draws
exactly as oracle/gen_golden.py does for the committed fixtures -- a guard against fixtures and
oracle drifting apart together."""
import numpy as np
import pytest

from oracle import philox, ref_shim
from oracle import pomdp_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference is not mounted here")


@pytest.fixture(scope="module")
def E():
    return ref_shim.load_reference()


def test_rock_steps_live(E):
    from gym_pomdp.envs.coord import Coord
    d = ref_shim.draws()
    rs = np.random.RandomState(0)
    with ref_shim.scripted_numpy():
        for n, k, stoch in [(7, 8, False), (11, 11, False), (11, 11, True)]:
            env = E.StochasticRockEnv(n, k) if stoch else E.RockEnv(n, k)
            cfg = O.RockCfg(n, k, stoch)
            rock_pos = [tuple(c) for c in env._rock_pos]
            for i in range(300):
                x, y = int(rs.randint(n)), int(rs.randint(n))
                st = rs.randint(-1, 2, k).tolist()
                a = int(rs.randint(5 + k))
                w = philox.draw_slots(99, np.array([i]), 1, philox.DOMAIN_STEP, 2)[0]
                d.clear(); d.feed([0] * k)
                env._set_state({"agent_pos": (x, y), "target": -1,
                                "rocks": [{"status": s, "pos": Coord(*rock_pos[j]), "count": 0, "measured": 0, "lkw": 1., "lkv": 1.,
                                           "prob_valuable": .5} for j, s in enumerate(st)]})
                d.clear(); d.feed([w[0], w[1]] if stoch else [w[1]])
                ob, rw, done, _ = env.step(a)
                ex, ey, est, eob, erw, edone, err = O.rock_step(cfg, x, y, st, a, lambda s: int(w[s]))
                assert (env.state.agent_pos.x, env.state.agent_pos.y, [r.status for r in env.state.rocks], ob, rw, bool(done)) == \
                    (ex, ey, est, eob, erw, edone) and err == 0
                assert [int(v) for v in env._generate_legal()] == O.rock_generate_legal(cfg, ex, ey, est)
        d.clear()


def test_tiger_and_network_steps_live(E):
    d = ref_shim.draws()
    rs = np.random.RandomState(1)
    with ref_shim.scripted_numpy():
        env = E.TigerEnv()
        d.feed_gym([0]); env.reset()
        for i in range(300):
            s, a = int(rs.randint(2)), int(rs.randint(3))
            w = philox.draw_slots(98, np.array([i]), 1, philox.DOMAIN_STEP, 1)[0]
            env._set_state(s)
            d.clear(); d.feed_gym([w[0]]); d.feed([w[0]])            # ONE word serves state_space.sample() and uniform()
            ob, rw, done, _ = env.step(a)
            assert (env.state, ob, rw, bool(done)) == O.tiger_step(s, a, lambda k: int(w[k]))
        env = E.NetworkEnv()
        nb = O.network_neighbours(10, 3)
        assert [list(map(int, r)) for r in env.neighbours] == nb
        for i in range(300):
            s, a = int(rs.randint(1024)), int(rs.randint(21))
            bits = np.array([(s >> m) & 1 for m in range(10)], np.int8)
            w = philox.draw_slots(97, np.array([i]), 1, philox.DOMAIN_STEP, 11)[0]
            env.reset(); env._set_state(bits.copy())
            d.clear(); d.feed([w[m] for m in range(10) if bits[m]] + ([w[10]] if a < 20 else []))
            ob, rw, done, info = env.step(a)
            es, eob, tenths, _ = O.network_step(bits.tolist(), a, lambda k: int(w[k]), nb)
            assert (list(map(int, info["state"])), int(ob), rw) == (es, eob, tenths / 10.0)
        d.clear()


def test_host_geometry_mirror_live(E):
    """gym_pomdp_b200.geometry's host classes next to the reference's coord.py / tag.py ones: board, indexing (including
    numpy's negative-index wrap and the out-of-range None), codecs, the three distances, directional_distance as written
    (coord.py:87-98) and the sampling helpers under the same numpy seed."""
    import gym_pomdp.envs.coord as RC
    from gym_pomdp.envs.tag import TagGrid as RefTagGrid
    from gym_pomdp_b200 import geometry as G
    for xs, ys in [(7, 7), (10, 5), (11, 11)]:
        a, b = RC.Grid(xs, ys), G.Grid(xs, ys)
        assert np.array_equal(a.board, b.board) and a.board.dtype == b.board.dtype and a.get_size == b.get_size
        a[RC.Coord(1, 2)] = 5
        b[G.Coord(1, 2)] = 5
        for c in [(1, 2), (0, 0), (-1, 0), (xs, 0), (0, ys), (xs - 1, ys - 1), (-xs, -ys), (-xs - 1, 0)]:
            ra, rb = a[RC.Coord(*c)], b[G.Coord(*c)]
            assert (ra is None and rb is None) or ra == rb, c
        assert [list(r) for r in a] == [list(r) for r in b]
        for i in range(xs * ys):
            assert tuple(a.get_coord(i)) == tuple(b.get_coord(i)) and a.get_index(a.get_coord(i)) == b.get_index(b.get_coord(i))
        rs = np.random.RandomState(3)
        for _ in range(200):
            c1, c2 = rs.randint(-2, 16, 2).tolist(), rs.randint(-2, 16, 2).tolist()
            assert a.is_inside(RC.Coord(*c1)) == b.is_inside(G.Coord(*c1))
            assert RC.Grid.euclidean_distance(RC.Coord(*c1), RC.Coord(*c2)) == G.Grid.euclidean_distance(G.Coord(*c1), G.Coord(*c2))
            assert RC.Grid.manhattan_distance(RC.Coord(*c1), RC.Coord(*c2)) == G.Grid.manhattan_distance(G.Coord(*c1), G.Coord(*c2))
            for d in range(4):
                assert RC.Grid.directional_distance(RC.Coord(*c1), RC.Coord(*c2), d) == \
                    G.Grid.directional_distance(G.Coord(*c1), G.Coord(*c2), d)
        with pytest.raises(NotImplementedError):
            G.Grid.directional_distance(G.Coord(0, 0), G.Coord(1, 1), 4)
        np.random.seed(7)
        ra = [tuple(a.sample()) for _ in range(20)] + [RC.Moves.sample() for _ in range(20)]
        np.random.seed(7)
        rb = [tuple(b.sample()) for _ in range(20)] + [G.Moves.sample() for _ in range(20)]
        assert ra == rb
    ta, tb = RefTagGrid((10, 5)), G.TagGrid((10, 5))
    np.random.seed(9)
    sa = [tuple(ta.sample()) for _ in range(50)]
    np.random.seed(9)
    assert sa == [tuple(tb.sample()) for _ in range(50)]
    assert np.array_equal(ta.board, tb.board)
