"""Record OUTCOME COUNTS of the unmodified reference's stochastic transitions, drawn with the
reference's own RNG (numpy's global MT19937, seeded below; gym's ``Discrete.sample`` mapped
to ``np.random.randint`` as real gym does), into tests/golden/ref_dist.npz.

TEST INFRASTRUCTURE ONLY; runs in the build container (the GPU box has no /root/reference).
Run:  python oracle/gen_dist.py        (about 3 minutes, deterministic)

The parity fixtures of gen_golden.py couple the reference to Philox words draw by draw.
This file is the complementary, UNCOUPLED pin: nothing here knows about Philox, the draw
slots or the threshold arithmetic, so the two-sample chi-square tests in
tests/test_stochastic_dist.py compare the kernels' per-action outcome distributions (10^7
draws) with what the reference itself does when left alone (north_star: "match its
per-action distribution to chi-sq p>0.01 on 10^7 draws under a stated seed").
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_dist.npz")
NP_SEED = 20261018   # stated seed; every env section re-seeds with NP_SEED + section index (TigerEnv.__init__
                     # itself calls np.random.seed(0), tiger.py:58, which would otherwise pin everything after it)

# (x, y, rock, status) probes for the Rock(11,11) sensor: L1 distances 0, 1, 4, 7, 12, 20
ROCK_CASES = [(0, 3, 0, 1), (0, 4, 0, -1), (3, 4, 0, 1), (2, 8, 0, -1), (9, 6, 0, 1), (10, 10, 1, -1), (10, 0, 2, 1),
              (5, 5, 10, 0)]
# (agent cell, opponent cell) probes for Tag's failed-TAG opponent move
TAG_CASES = [(13, 15), (24, 27), (0, 19), (15, 13), (22, 25), (9, 0), (27, 21), (5, 26)]
NETWORK_STATES = [0b1111111111, 0b1111111110, 0b1011011101, 0b0000000001]


def main():
    E = ref_shim.load_reference()
    import gym
    from gym_pomdp.envs.coord import Coord
    from gym_pomdp.envs.tag import TagState
    gym.spaces.Discrete.sample = lambda self: int(np.random.randint(self.n))    # what gym's Discrete does
    warnings.simplefilter("ignore")            # rock.py:191 divides 0/0 once lkv/lkw underflow (reference behaviour)
    np.random.seed(NP_SEED)
    out = {"np_seed": NP_SEED}

    # ---- RockSample(11,11): check action, obs counts [BAD=1, GOOD=2]  (rock.py:171-175, 401-407)
    T = 100000
    env = E.RockEnv(11, 11)
    np.random.seed(NP_SEED + 1)
    env.reset()
    counts = np.zeros((len(ROCK_CASES), 3), np.int64)
    for c, (x, y, rock, status) in enumerate(ROCK_CASES):
        env.state.agent_pos = Coord(x, y)
        env.state.rocks[rock].status = status
        for _ in range(T):
            ob, rw, done, _ = env.step(5 + rock)
            counts[c, ob] += 1
        assert not done and rw == 0
    out["rock_cases"] = np.array(ROCK_CASES, np.int32)
    out["rock_obs_counts"] = counts
    print("rock sensor", counts.tolist())

    # ---- RockSample reset: how often each rock starts good (rock.py:78-80)
    env = E.RockEnv(7, 8)
    np.random.seed(NP_SEED + 2)
    good = np.zeros(8, np.int64)
    R = 50000
    for _ in range(R):
        env.reset()
        good += [r.status == 1 for r in env.state.rocks]
    out["rock_reset_good"] = good
    out["rock_reset_trials"] = R

    # ---- StochasticRock(7,8): the p_move gate on a NORTH move from (3,3)  (rock.py:443)
    env = E.StochasticRockEnv(7, 8)
    np.random.seed(NP_SEED + 3)
    env.reset()
    moved = 0
    for _ in range(T):
        env.state.agent_pos = Coord(3, 3)
        env.step(0)
        moved += env.state.agent_pos.y == 4
    out["srock_moved"] = np.array([T - moved, moved], np.int64)
    print("stochastic rock moved", moved / T)

    # ---- Tag: opponent cell after a failed TAG  (tag.py:119-131, 201-207, 260-280)
    env = E.TagEnv()
    np.random.seed(NP_SEED + 4)
    env.reset()
    g = env.grid
    tc = np.zeros((len(TAG_CASES), 29), np.int64)
    for c, (a, o) in enumerate(TAG_CASES):
        for _ in range(T):
            st = TagState(g.get_tag_coord(a))
            st.opponent_pos = [g.get_tag_coord(o)]
            st.num_opp = 1
            env._set_state(st)
            ob, rw, done, _ = env.step(4)
            tc[c, g.get_index(env.state.opponent_pos[0])] += 1
        assert rw == -10.
    out["tag_cases"] = np.array(TAG_CASES, np.int32)
    out["tag_opp_counts"] = tc
    print("tag", [row[row > 0].tolist() for row in tc])
    # Tag reset: agent / opponent cells and the reset observation (tag.py:97-102, 181-193)
    ra, ro, rob = np.zeros(29, np.int64), np.zeros(29, np.int64), np.zeros(30, np.int64)
    for _ in range(T):
        ob = env.reset()
        ra[g.get_index(env.state.agent_pos)] += 1
        ro[g.get_index(env.state.opponent_pos[0])] += 1
        rob[ob] += 1
    out.update(tag_reset_agent=ra, tag_reset_opp=ro, tag_reset_obs=rob)

    # ---- Tiger: listen observation per state; state after opening the safe door  (tiger.py:72-88, 140-149)
    env = E.TigerEnv()
    np.random.seed(NP_SEED + 5)
    env.reset()
    listen = np.zeros((2, 3), np.int64)
    for s in (0, 1):
        for _ in range(T):
            env._set_state(s)
            ob, rw, done, _ = env.step(2)
            listen[s, ob] += 1
    resample = np.zeros((2, 2), np.int64)
    for s in (0, 1):
        for _ in range(T):
            env._set_state(s)
            ob, rw, done, _ = env.step(1 - s)
            assert ob == 2 and rw == 10 and not done
            resample[s, env.state] += 1
    rst = np.zeros(2, np.int64)
    for _ in range(T):
        env.reset()
        rst[env.state] += 1
    out.update(tiger_listen=listen, tiger_resample=resample, tiger_reset=rst)
    print("tiger listen", listen.tolist(), "resample", resample.tolist())

    # ---- Network: machines still up after one step; ping / reboot observations  (network.py:71-114)
    env = E.NetworkEnv()
    np.random.seed(NP_SEED + 6)
    env.reset()
    up = np.zeros((len(NETWORK_STATES), 10), np.int64)
    for c, s in enumerate(NETWORK_STATES):
        bits = np.array([(s >> m) & 1 for m in range(10)], np.int8)
        for _ in range(T):
            env._set_state(bits.copy())
            env.step(20)
            up[c] += env.state
    ping = np.zeros((2, 3), np.int64)            # [post-state bit of machine 1][obs]
    reboot = np.zeros(3, np.int64)
    for _ in range(T):
        env._set_state(np.ones(10, np.int8))
        ob, _, _, _ = env.step(2)                # ping machine 1
        ping[env.state[1], ob] += 1
        env._set_state(np.ones(10, np.int8))
        ob, _, _, _ = env.step(3)                # reboot machine 1
        reboot[ob] += 1
    out.update(network_states=np.array(NETWORK_STATES, np.int64), network_up=up, network_ping=ping,
               network_reboot=reboot, trials=T)
    print("network up", (up / T).round(3).tolist())

    # ---- BattleShip: where the first (length 3) and second (length 2) ship land  (battleship.py:167-211)
    for xs, ys, R in [(5, 5, 40000), (10, 10, 40000)]:
        env = E.BattleShipEnv(board_size=(xs, ys))
        np.random.seed(NP_SEED + 7 + xs)
        first = np.zeros(4 * xs * ys, np.int64)
        cells = np.zeros((xs, ys), np.int64)
        for _ in range(R):
            env.reset()
            sh = env.state.ships[0]           # Ship.pos stays the start cell: mark_ship rebinds a local (battleship.py:184-193)
            cells += np.array([[env.grid.board[x, y].occupied for y in range(ys)] for x in range(xs)])
            first[4 * env.grid.get_index(sh.pos) + sh.direction] += 1
        out[f"ship_{xs}x{ys}_first"] = first
        out[f"ship_{xs}x{ys}_cells"] = cells
        out[f"ship_{xs}x{ys}_trials"] = R
        print(f"battleship {xs}x{ys}: first-ship placements seen {int((first > 0).sum())}")

    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
