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
            w = philox.draw_slots(98, np.array([i]), 1, philox.DOMAIN_STEP, 2)[0]
            env._set_state(s)
            d.clear(); d.feed_gym([w[0]]); d.feed([w[1]])
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
