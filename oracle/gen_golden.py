"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY; runs in the build container (the GPU box has no /root/reference).
Run:  python oracle/gen_golden.py           (rewrites every fixture, deterministic)

The reference ships no golden vectors (six Coord asserts at coord.py:121-126
are all there is), so the fixtures are outputs of the reference itself, with its
np.random / gym draws scripted from Philox words through oracle/ref_shim.py.  All arrays
are in the reference's own units (coords, statuses in {-1,0,+1}, float64 rewards); nothing
here knows about the packed device layout.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import philox, ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SEED = 0x5EED


def words(n_cases, n_slots, stream):
    """Philox draw words [n_cases, n_slots] for fixture stream ``stream`` (used as step)."""
    return philox.draw_slots(SEED, np.arange(n_cases), stream, philox.DOMAIN_STEP, n_slots)


def rng_ints(n_cases, hi, stream):
    """Deterministic uniform ints in [0, hi) for building synthetic inputs."""
    w = philox.draw_slots(SEED ^ 0xABCDEF, np.arange(n_cases), stream, 7, 1)[:, 0].astype(np.uint64)
    return ((w * np.uint64(hi)) >> np.uint64(32)).astype(np.int64)


# ---------------------------------------------------------------------------------------
def gen_coord(E):
    from gym_pomdp.envs.coord import Coord, Grid, Moves
    from gym_pomdp.envs.tag import TagGrid
    out = {}
    # the reference's only shipped KATs (coord.py:121-126)
    kat_in = [((3, 3), (2, 2)), ((5, 2), (2, 5)), ((2, 2), tuple(Moves.NORTH.value)),
              ((2, 2), tuple(Moves.WEST.value)), ((2, 2), tuple(Moves.SOUTH.value)),
              ((2, 2), tuple(Moves.EAST.value))]
    out["kat_a"] = np.array([a for a, _ in kat_in], dtype=np.int32)
    out["kat_b"] = np.array([b for _, b in kat_in], dtype=np.int32)
    out["kat_sum"] = np.array([tuple(Coord(*a) + Coord(*b)) for a, b in kat_in], dtype=np.int32)
    out["moves"] = np.array([tuple(Moves.get_coord(i)) for i in range(5)], dtype=np.int32)
    out["opposite"] = np.array([Grid.opposite(m) for m in range(4)], dtype=np.int32)
    for (xs, ys) in [(7, 7), (11, 11), (15, 15), (10, 10), (5, 5), (10, 5)]:
        g = Grid(xs, ys)
        idx = np.arange(g.n_tiles)
        coords = np.array([tuple(g.get_coord(i)) for i in idx], dtype=np.int32)
        back = np.array([g.get_index(Coord(*c)) for c in coords], dtype=np.int32)
        assert (back == idx).all()
        out[f"grid_{xs}x{ys}_coord"] = coords
        probe = np.array([(x, y) for x in range(-2, xs + 2) for y in range(-2, ys + 2)], dtype=np.int32)
        out[f"grid_{xs}x{ys}_probe"] = probe
        out[f"grid_{xs}x{ys}_inside"] = np.array([bool(g.is_inside(Coord(*c))) for c in probe])
    tg = TagGrid((10, 5), obs_cells=29)
    tc = np.array([tuple(tg.get_tag_coord(i)) for i in range(29)], dtype=np.int32)
    out["tag_coord"] = tc
    out["tag_index"] = np.array([tg.get_index(Coord(*c)) for c in tc], dtype=np.int32)
    probe = np.array([(x, y) for x in range(-2, 12) for y in range(-2, 7)], dtype=np.int32)
    out["tag_probe"] = probe
    out["tag_inside"] = np.array([bool(tg.is_inside(Coord(*c))) for c in probe])
    out["tag_corner"] = np.array([bool(tg.is_corner(Coord(*c))) for c in probe])
    # L1 "euclidean" distance table (coord.py:79-81)
    pts = np.array([(x, y) for x in range(0, 15, 2) for y in range(0, 15, 3)], dtype=np.int32)
    out["dist_pts"] = pts
    out["dist_l1"] = np.array([[Grid.euclidean_distance(Coord(*a), Coord(*b)) for b in pts] for a in pts])
    np.savez_compressed(os.path.join(GOLDEN, "coord.npz"), **out)
    print("coord.npz", {k: v.shape for k, v in list(out.items())[:4]}, "...")


# ---------------------------------------------------------------------------------------
def rock_state_dict(Coord, x, y, status, rock_pos):
    return {"agent_pos": (int(x), int(y)), "target": -1,
            "rocks": [{"status": int(s), "pos": Coord(*rock_pos[i]), "count": 0, "measured": 0,
                       "lkw": 1., "lkv": 1., "prob_valuable": .5} for i, s in enumerate(status)]}


def gen_rock(E, n, k, stochastic, n_random, tag):
    from gym_pomdp.envs.coord import Coord
    d = ref_shim.draws()
    env = E.StochasticRockEnv(n, k) if stochastic else E.RockEnv(n, k)
    rock_pos = [tuple(c) for c in env._rock_pos]
    A = env.action_space.n
    max_d = 2 * (n - 1)
    out = {"n": n, "k": k, "stochastic": stochastic, "n_actions": A,
           "grid": env.grid.board.astype(np.int8).copy(),  # [x, y]
           "rock_pos": np.array(rock_pos, dtype=np.int32),
           "start": np.array(tuple(env._agent_pos), dtype=np.int32),
           "eff": np.array([E.RockEnv._efficiency(Coord(0, 0), Coord(dd, 0)) for dd in range(max_d + 1)]),
           "discount": env._discount, "reward_range": env._reward_range}
    # ---- inputs: exhaustive deterministic part + random full states
    xs, ys, sts, acts = [], [], [], []
    for x in range(n):
        for y in range(n):
            for s_under in (-1, 0, 1):
                for a in range(5):
                    st = [(-1, 0, 1)[(x + 2 * y + i + a) % 3] for i in range(k)]
                    g = env.grid.board[x, y]
                    if 0 <= g < k:
                        st[g] = s_under
                    xs.append(x); ys.append(y); sts.append(st); acts.append(a)
    n_ex = len(xs)
    rx = rng_ints(n_random, n, 1); ry = rng_ints(n_random, n, 2); ra = rng_ints(n_random, A, 3)
    for i in range(n_random):
        st = (rng_ints(k, 3, 100 + i) - 1).tolist()
        xs.append(int(rx[i])); ys.append(int(ry[i])); sts.append(st); acts.append(int(ra[i]))
    N = len(xs)
    dr = words(N, 2, stream=11)
    # boundary draws for the sensor: exactly at the Bernoulli threshold and one below
    for i in range(n_ex, min(N, n_ex + 400)):
        if acts[i] > 4:
            dd = abs(xs[i] - rock_pos[acts[i] - 5][0]) + abs(ys[i] - rock_pos[acts[i] - 5][1])
            T = int(np.ceil(out["eff"][dd] * 2.0 ** 32))
            dr[i, 1] = min(T - (i & 1), 0xFFFFFFFF)
    res = {k_: [] for k_ in ("x2", "y2", "st2", "ob", "rw", "done", "raised", "ndraw", "prob", "legal")}
    for i in range(N):
        d.clear(); d.feed([0] * k)
        env._set_state(rock_state_dict(Coord, xs[i], ys[i], sts[i], rock_pos))
        d.clear(); d.consumed = 0
        d.feed([dr[i, 0], dr[i, 1]] if stochastic else [dr[i, 1]])
        raised = False
        try:
            ob, rw, done, info = env.step(acts[i])
        except IndexError:
            raised, ob, rw, done = True, 0, 0, False
        res["ndraw"].append(d.consumed)
        res["raised"].append(raised)
        res["x2"].append(env.state.agent_pos.x); res["y2"].append(env.state.agent_pos.y)
        res["st2"].append([r.status for r in env.state.rocks])
        res["ob"].append(ob); res["rw"].append(rw); res["done"].append(bool(done))
        if raised:
            res["prob"].append([0.] * 3); res["legal"].append([-1] * (A + 4))
            continue
        # NEXT rows: observation likelihood and legal set on the post state
        post = rock_state_dict(Coord, env.state.agent_pos.x, env.state.agent_pos.y,
                               [r.status for r in env.state.rocks], rock_pos)
        pr = []
        for o in range(3):
            d.clear(); d.feed([0] * k)
            pr.append(float(env._compute_prob(acts[i], post, o)))
        res["prob"].append(pr)
        try:
            lg = [int(a) for a in env._generate_legal()]
        except (IndexError, AssertionError):
            lg = []
        res["legal"].append(lg + [-1] * (A + 4 - len(lg)))
    out.update(x=np.array(xs, np.int32), y=np.array(ys, np.int32), status=np.array(sts, np.int8),
               action=np.array(acts, np.int32), draws=dr, n_exhaustive=n_ex,
               x2=np.array(res["x2"], np.int32), y2=np.array(res["y2"], np.int32),
               status2=np.array(res["st2"], np.int8), obs=np.array(res["ob"], np.int32),
               reward=np.array(res["rw"], np.float64), done=np.array(res["done"]),
               raised=np.array(res["raised"]), ndraw=np.array(res["ndraw"], np.int32),
               prob=np.array(res["prob"], np.float64), legal=np.array(res["legal"], np.int8))
    # ---- reset (rock.py:236-241): rock i's uniform = rotl32(word of slot 0, 30 - 2 i) / 2^32
    M = 512
    rd = philox.draw_slots(SEED, np.arange(M), 0, philox.DOMAIN_RESET, 1)
    rd[0, 0] = 1 << 1                  # rock 0: u == 0.5 exactly -> sign() gives 0
    rd[1, 0] = (1 << 31) - 1
    rd[2, 0] = (1 << 31) + 1
    rd[3, 0] = 1 << (2 * (k - 1) + 1)  # the last rock exactly at one half
    rd[4, 0] = 0                       # every rock below one half
    rd[5, 0] = 0xFFFFFFFF              # every rock above
    rd[6, 0] = 1                       # a single bit that decides nothing: nobody ties
    rd[7, 0] = 0xAAAAAAAA
    rd[8, 0] = 0x55555555
    rd[9, 0] = 1 << 9                  # rock 4 ties
    rst, rob, rxy = [], [], []
    from oracle.pomdp_oracle import rock_reset_word
    for i in range(M):
        d.clear(); d.feed([rock_reset_word(int(rd[i, 0]), r) for r in range(k)])
        rob.append(env.reset())
        rst.append([r.status for r in env.state.rocks]); rxy.append(tuple(env.state.agent_pos))
        assert not d.np_queue
    out.update(reset_draws=rd, reset_status=np.array(rst, np.int8), reset_obs=np.array(rob, np.int32),
               reset_xy=np.array(rxy, np.int32))
    np.savez_compressed(os.path.join(GOLDEN, f"rock_{tag}.npz"), **out)
    print(f"rock_{tag}.npz cases={N} raised={int(np.sum(res['raised']))} done={int(np.sum(res['done']))}")


# ---------------------------------------------------------------------------------------
def gen_tag(E, n_opp, tag):
    from gym_pomdp.envs.coord import Coord
    from gym_pomdp.envs.tag import TagState
    from oracle.pomdp_oracle import tag_pick_word
    d = ref_shim.draws()
    env = E.TagEnv(num_opponents=n_opp)
    g = env.grid
    out = {"n_opp": n_opp, "n_actions": env.action_space.n, "n_obs": env.observation_space.n,
           "move_prob": env.move_prob, "discount": env._discount, "reward_range": env._reward_range}
    # admissible-move multisets for all 29x29 (agent, opp) pairs (tag.py:260-280)
    from gym_pomdp.envs.coord import Moves
    order = list(Moves)
    adm = -np.ones((29, 29, 4), np.int8)
    for a in range(29):
        for o in range(29):
            acts = E.TagEnv._admissable_actions(g.get_tag_coord(a), g.get_tag_coord(o)) if a != o else []
            if a == o:
                acts = E.TagEnv._admissable_actions(g.get_tag_coord(a), g.get_tag_coord(o))
            for j, m in enumerate(acts):
                adm[a, o, j] = order.index(m)
    out["admissible"] = adm
    ag, ops, nop, acts = [], [], [], []
    if n_opp == 1:
        for a in range(29):
            for o in range(29):
                for act in range(5):
                    ag.append(a); ops.append([o]); nop.append(1); acts.append(act)
    n_ex = len(ag)
    R = 3000
    ra = rng_ints(R, 29, 21); ract = rng_ints(R, 5, 22)
    for i in range(R):
        o = rng_ints(n_opp, 29, 300 + i).tolist()
        if i % 3 == 0:
            o[0] = int(ra[i])  # make tags / "29" observations common
        ag.append(int(ra[i])); ops.append(o); acts.append(int(ract[i] if i % 2 else 4))
        nop.append(n_opp if n_opp == 1 else 1 + int(rng_ints(1, n_opp, 900 + i)[0]))
    N = len(ag)
    dr = words(N, n_opp, stream=12)
    T = int(np.ceil(0.8 * 2.0 ** 32))
    for i in range(n_ex, min(N, n_ex + 200)):
        dr[i, 0] = T - (i & 1)  # Bernoulli(0.8) boundary
    # map the reference's sequential draws onto slots: wrap move_opponent (test harness only)
    orig_move = env.move_opponent
    cur = {"i": 0}

    def move_opponent(opp):
        d.clear(); d.feed([dr[cur["i"], opp], tag_pick_word(dr[cur["i"], opp])])
        orig_move(opp)
        d.clear()
    env.move_opponent = move_opponent
    res = {k_: [] for k_ in ("ag2", "op2", "nop2", "ob", "rw", "done", "prob")}
    for i in range(N):
        st = TagState(g.get_tag_coord(ag[i]))
        st.opponent_pos = [g.get_tag_coord(o) for o in ops[i]]
        st.num_opp = nop[i]
        env._set_state(st)
        cur["i"] = i
        ob, rw, done, info = env.step(acts[i])
        res["ag2"].append(g.get_index(env.state.agent_pos))
        res["op2"].append([g.get_index(o) for o in env.state.opponent_pos])
        res["nop2"].append(env.state.num_opp)
        res["ob"].append(ob); res["rw"].append(rw); res["done"].append(bool(done))
        res["prob"].append([float(env._compute_prob(acts[i], env.state, o)) for o in range(30)])
    out.update(agent=np.array(ag, np.int32), opp=np.array(ops, np.int32), num_opp=np.array(nop, np.int32),
               action=np.array(acts, np.int32), draws=dr, n_exhaustive=n_ex,
               agent2=np.array(res["ag2"], np.int32), opp2=np.array(res["op2"], np.int32),
               num_opp2=np.array(res["nop2"], np.int32), obs=np.array(res["ob"], np.int32),
               reward=np.array(res["rw"], np.float64), done=np.array(res["done"]),
               prob=np.array(res["prob"], np.float64))
    # reset (tag.py:97-102): the j-th randint(29) (agent, then each opponent) = digit j % 3 of slot j // 3
    from oracle.pomdp_oracle import tag_reset_word
    env.move_opponent = orig_move
    M = 600
    n_cells = 1 + n_opp
    rd = philox.draw_slots(SEED, np.arange(M), 0, philox.DOMAIN_RESET, (n_cells + 2) // 3)
    for i in range(64):                # coinciding agent / opponent 0 -> reset ob 29: digits (a, a, i % 29) of slot 0
        a = i % 29
        rd[i, 0] = -(-(((a * 29 + a) * 29 + i % 29) * 2 ** 32 + 2 ** 31) // 29 ** 3)
    rob, rag, rop = [], [], []
    for i in range(M):
        d.clear(); d.feed([tag_reset_word(int(rd[i, j // 3]), j % 3) for j in range(n_cells)])
        rob.append(env.reset())
        rag.append(g.get_index(env.state.agent_pos)); rop.append([g.get_index(o) for o in env.state.opponent_pos])
        assert not d.np_queue and env.state.num_opp == n_opp
    out.update(reset_draws=rd, reset_obs=np.array(rob, np.int32), reset_agent=np.array(rag, np.int32),
               reset_opp=np.array(rop, np.int32))
    np.savez_compressed(os.path.join(GOLDEN, f"tag_{tag}.npz"), **out)
    print(f"tag_{tag}.npz cases={N} done={int(np.sum(res['done']))} ob29={int(np.sum(np.array(res['ob']) == 29))}")


# ---------------------------------------------------------------------------------------
def gen_battleship(E, xs, ys, max_len, tag, n_boards):
    from gym_pomdp.envs.battleship import Ship, ShipState
    from gym_pomdp.envs.coord import Coord
    d = ref_shim.draws()
    env = E.BattleShipEnv(board_size=(xs, ys), max_len=max_len)
    nt = xs * ys
    out = {"x_size": xs, "y_size": ys, "max_len": max_len, "n_actions": env.action_space.n,
           "discount": env._discount, "reward_range": env._reward_range}
    ATT = 48 if xs * ys >= 100 else 320
    rd = philox.draw_slots(SEED, np.arange(n_boards), 0, philox.DOMAIN_RESET, 2 * ATT)
    occ = np.zeros((n_boards, xs, ys), bool)
    attempts = np.zeros(n_boards, np.int32)
    ships = np.zeros((n_boards, max_len - 1, 4), np.int32)  # x, y, dir, length
    vis_in = np.zeros((n_boards, xs, ys), bool)
    S = 6  # steps recorded per board
    act = np.zeros((n_boards, S), np.int32)
    ob = np.zeros((n_boards, S), np.int32); rw = np.zeros((n_boards, S)); dn = np.zeros((n_boards, S), bool)
    rem = np.zeros((n_boards, S + 1), np.int32)
    prob = np.zeros((n_boards, S, 2)); legal_cnt = np.zeros((n_boards, S), np.int32)
    valid2 = np.zeros((n_boards, nt * 4), bool)  # candidates for the 2nd ship given the 1st
    lengths = list(reversed(range(2, max_len + 1)))
    for b in range(n_boards):
        d.clear(); d.consumed = 0; d.feed(rd[b])
        assert env.reset() == 0
        attempts[b] = d.consumed // 2
        d.clear()
        for x in range(xs):
            for y in range(ys):
                occ[b, x, y] = env.grid[Coord(x, y)].occupied
        for s, sh in enumerate(env.state.ships):
            ships[b, s] = (sh.pos.x, sh.pos.y, sh.direction, sh.length)
        # valid second-ship placements given the first ship only (battleship.py:195-211)
        if len(lengths) > 1:
            g2 = type(env.grid)((xs, ys)); st2 = ShipState()
            d.feed([0]); first = Ship(Coord(int(ships[b, 0, 0]), int(ships[b, 0, 1])), int(ships[b, 0, 3]))
            first.direction = int(ships[b, 0, 2])
            E.BattleShipEnv.mark_ship(first, g2, st2)
            for pos in range(nt):
                for dd in range(4):
                    d.clear(); d.feed([0])
                    sh = Ship(g2.get_coord(pos), lengths[1]); sh.direction = dd
                    valid2[b, 4 * pos + dd] = not E.BattleShipEnv.collision(sh, g2, st2)
            d.clear()
        # synthetic visited pattern (Bernoulli(0.3)), total_remaining recomputed
        v = philox.draw_slots(SEED ^ 0x51, np.arange(nt), b, 7, 1)[:, 0] < np.uint32(0.3 * 2 ** 32)
        v = v.reshape(xs, ys)
        if b % 4 == 0:
            v[:] = False
        remaining = int(np.sum(occ[b] & ~v))
        if remaining == 0:
            v[:] = False; remaining = int(occ[b].sum())
        vis_in[b] = v
        for x in range(xs):
            for y in range(ys):
                env.grid[Coord(x, y)].visited = bool(v[x, y])
        env.state.total_remaining = remaining
        rem[b, 0] = remaining
        a_seq = rng_ints(S, nt, 5000 + b)
        if b % 5 == 0:  # force hits so that wins (+n_tiles) occur in the fixture
            hits = [(y * xs + x) for x in range(xs) for y in range(ys) if occ[b, x, y] and not v[x, y]]
            a_seq = np.array((hits + hits)[:S] if len(hits) >= 1 else a_seq)
            if len(a_seq) < S:
                a_seq = np.resize(a_seq, S)
        for s in range(S):
            if env.done:
                act[b, s:] = -1
                rem[b, s + 1:] = env.state.total_remaining
                break
            act[b, s] = a_seq[s]
            ob[b, s], rw[b, s], dn[b, s], _ = env.step(int(a_seq[s]))
            rem[b, s + 1] = env.state.total_remaining
            prob[b, s] = [env._compute_prob(int(a_seq[s]), env.state, o) for o in range(2)]
            legal_cnt[b, s] = len(env._generate_legal())
    # empty-board candidates for the first ship
    g0 = type(env.grid)((xs, ys)); st0 = ShipState()
    valid1 = np.zeros(nt * 4, bool)
    for pos in range(nt):
        for dd in range(4):
            d.clear(); d.feed([0])
            sh = Ship(g0.get_coord(pos), lengths[0]); sh.direction = dd
            valid1[4 * pos + dd] = not E.BattleShipEnv.collision(sh, g0, st0)
    d.clear()
    out.update(reset_draws=rd, occupied=occ, attempts=attempts, ships=ships, visited_in=vis_in,
               action=act, obs=ob, reward=rw, done=dn, remaining=rem, prob=prob, legal_count=legal_cnt,
               valid_first=valid1, valid_second=valid2)
    np.savez_compressed(os.path.join(GOLDEN, f"battleship_{tag}.npz"), **out)
    print(f"battleship_{tag}.npz boards={n_boards} valid_first={int(valid1.sum())}/{nt * 4} "
          f"max_attempts={int(attempts.max())} wins={int(dn.sum())}")


# ---------------------------------------------------------------------------------------
def gen_tiger(E):
    d = ref_shim.draws()
    env = E.TigerEnv()
    d.feed_gym([0]); env.reset()  # step() needs the counters reset() creates (tiger.py:62)
    R = 300
    G = int(np.floor(0.85 * 2.0 ** 32))
    st, ac, dr = [], [], []
    w = words(6 * R, 1, stream=13)
    i = 0
    for s in range(2):
        for a in range(3):
            for j in range(R):
                ww = w[i].copy()
                if j < 4:
                    ww[0] = G + (j - 1)  # around the ``p > .85`` boundary
                st.append(s); ac.append(a); dr.append(ww); i += 1
    dr = np.array(dr, np.uint32)
    res = {k_: [] for k_ in ("s2", "ob", "rw", "done", "prob")}
    for i in range(len(st)):
        env._set_state(st[i])
        d.clear(); d.feed_gym([dr[i, 0]]); d.feed([dr[i, 0]])        # ONE word serves sample() and uniform()
        ob, rw, done, info = env.step(ac[i])
        res["s2"].append(env.state); res["ob"].append(ob); res["rw"].append(rw); res["done"].append(bool(done))
        res["prob"].append([float(E.TigerEnv._compute_prob(ac[i], env.state, o)) for o in range(3)])
    rd = philox.draw_slots(SEED, np.arange(256), 0, philox.DOMAIN_RESET, 1)
    rs, ro = [], []
    for i in range(256):
        d.clear(); d.feed_gym(rd[i]); ro.append(env.reset()); rs.append(env.state)
    d.clear()
    np.savez_compressed(os.path.join(GOLDEN, "tiger.npz"), state=np.array(st, np.int32),
                        action=np.array(ac, np.int32), draws=dr, state2=np.array(res["s2"], np.int32),
                        obs=np.array(res["ob"], np.int32), reward=np.array(res["rw"], np.float64),
                        done=np.array(res["done"]), prob=np.array(res["prob"]), reset_draws=rd,
                        reset_state=np.array(rs, np.int32), reset_obs=np.array(ro, np.int32),
                        discount=env._discount, reward_range=env._reward_range)
    print("tiger.npz cases", len(st))


# ---------------------------------------------------------------------------------------
def gen_network(E, n, ptype, tag, exhaustive):
    d = ref_shim.draws()
    env = E.NetworkEnv(n_machines=n, problem_type=ptype)
    A = env.action_space.n
    nb = env.neighbours
    nb_arr = -np.ones((n, 3), np.int32)
    for i, l in enumerate(nb):
        nb_arr[i, :len(l)] = l
    sts, acts = [], []
    if exhaustive:
        for s in range(1 << n):
            for a in range(A):
                sts.append(s); acts.append(a)
    n_ex = len(sts)
    R = 4000
    rs = rng_ints(R, 1 << n, 31); ra = rng_ints(R, A, 32)
    sts += [int(v) for v in rs]; acts += [int(v) for v in ra]
    N = len(sts)
    # the per-machine words of the joint failure draw (slot g = machines 5g..5g+4) + the observation draw's word
    dr = philox.network_draws(SEED, np.arange(N), 14, n, env._p, env._q)
    s2, ob, rw, prob = [], [], [], []
    for i in range(N):
        bits = np.array([(sts[i] >> m) & 1 for m in range(n)], np.int8)
        env.reset(); env._set_state(bits.copy())
        order = [m for m in range(n) if bits[m]] + ([n] if acts[i] < 2 * n else [])
        d.clear(); d.feed([dr[i, m] for m in order])
        o, r, done, info = env.step(acts[i])
        assert not d.np_queue and done is False
        s2.append(int(sum(int(v) << m for m, v in enumerate(info["state"]))))
        ob.append(int(o)); rw.append(float(r))
        prob.append([float(env._compute_prob(acts[i], info["state"], oo)) for oo in range(3)])
    env.reset()
    np.savez_compressed(os.path.join(GOLDEN, f"network_{tag}.npz"), n=n, problem_type=ptype, n_actions=A,
                        neighbours=nb_arr, state=np.array(sts, np.int64), action=np.array(acts, np.int32),
                        draws=dr, n_exhaustive=n_ex, state2=np.array(s2, np.int64), obs=np.array(ob, np.int32),
                        reward=np.array(rw, np.float64), prob=np.array(prob), reset_obs=0,
                        reset_state=int(sum(int(v) << m for m, v in enumerate(env.state))),
                        discount=env._discount, reward_range=env._reward_range, p=env._p, q=env._q, p_ob=env._p_ob)
    print(f"network_{tag}.npz cases={N}")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    assert philox.kat()
    E = ref_shim.load_reference()
    with ref_shim.scripted_numpy():
        gen_coord(E)
        gen_rock(E, 7, 8, False, 3000, "7_8")
        gen_rock(E, 11, 11, False, 4000, "11_11")
        gen_rock(E, 15, 15, False, 4000, "15_15")
        gen_rock(E, 7, 7, False, 1000, "7_7")
        gen_rock(E, 4, 3, False, 500, "4_3")
        gen_rock(E, 7, 8, True, 3000, "stoch_7_8")
        gen_rock(E, 11, 11, True, 2000, "stoch_11_11")
        gen_tag(E, 1, "1opp")
        gen_tag(E, 2, "2opp")
        gen_battleship(E, 10, 10, 3, "10x10", 160)
        gen_battleship(E, 5, 5, 3, "5x5", 120)
        gen_tiger(E)
        gen_network(E, 10, 3, "3legs10", True)
        gen_network(E, 7, 3, "3legs7", True)
        gen_network(E, 10, 1, "ring10", False)
        gen_network(E, 19, 3, "3legs19", False)


if __name__ == "__main__":
    main()
