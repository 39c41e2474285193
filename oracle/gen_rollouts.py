"""Record ROLLOUTS of the unmodified reference under its own ``_generate_legal`` + ``step``
(the loops at rock.py:563-572 and tag.py:310-316; rock.py:563's ``_generate_preferred(history)`` is
``_generate_legal()`` for the default ``use_heuristic=False``, rock.py:293-295) into tests/golden/rollouts.npz.

TEST INFRASTRUCTURE ONLY; runs in the build container (the GPU box has no /root/reference).
Run:  python oracle/gen_rollouts.py         (deterministic, about 10 s)

Episode e of an env is instance e of a batch: seed 0x5EED, reset at counter 1, then rollout
step t at counter 2 + t.  The reference's draws are scripted from the Philox words of that
(instance, counter):  np.random.choice(legal) <- (domain POLICY, slot 0); the step's own
draws <- (domain STEP, its slots) in the order the reference consumes them.  Recorded per
episode: the state after reset, the actions taken, the discounted return exactly as the
Python loop accumulates it (``r += rw * discount; discount *= env._discount``), the number
of steps, the final state and ``done``.  Nothing here knows the packed device layout.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import philox, ref_shim  # noqa: E402
from oracle.pomdp_oracle import rock_reset_word, tag_pick_word, tag_reset_word  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SEED = 0x5EED
RESET_CTR, FIRST_CTR = 1, 2


def W(e, ctr, domain, n_slots):
    return philox.draw_slots(SEED, np.array([e]), ctr, domain, n_slots)[0]


def run(env, d, e, T, n_step_slots, feed_step, snapshot, gamma=None, step_words=None):
    """One episode of the reference's loop.  feed_step(words, action) scripts the step's draws; step_words(e, ctr)
    replaces the plain slot words (Network: the per-machine words of the joint failure draw)."""
    gamma = env._discount if gamma is None else gamma
    r, disc, t, acts, done = 0.0, 1.0, 0, [], False
    while t < T and not done:
        ctr = FIRST_CTR + t
        legal = env._generate_legal()
        d.clear()
        d.feed([W(e, ctr, philox.DOMAIN_POLICY, 1)[0]])
        a = int(np.random.choice(legal))                       # scripted: legal[floor(u * len)]
        d.clear()
        feed_step(step_words(e, ctr) if step_words else W(e, ctr, philox.DOMAIN_STEP, n_step_slots), a)
        ob, rw, done, info = env.step(a)
        d.clear()
        r += rw * disc
        disc *= gamma
        acts.append(a)
        t += 1
    return r, t, acts, bool(done), snapshot()


def pad(rows, T):
    out = -np.ones((len(rows), T), np.int32)
    for i, row in enumerate(rows):
        out[i, :len(row)] = row
    return out


def gen_rock(E, out, tag, n, k, stochastic, M, T):
    d = ref_shim.draws()
    env = E.StochasticRockEnv(n, k) if stochastic else E.RockEnv(n, k)
    res = {key: [] for key in ("x0", "y0", "st0", "ret", "steps", "acts", "done", "x1", "y1", "st1")}

    def snap():
        return env.state.agent_pos.x, env.state.agent_pos.y, [r.status for r in env.state.rocks]
    for e in range(M):
        rw = W(e, RESET_CTR, philox.DOMAIN_RESET, 1)
        d.clear(); d.feed([rock_reset_word(int(rw[0]), r) for r in range(k)])   # rock r's uniform(0,1), rock.py:78-80
        env.reset()
        d.clear()
        x0, y0, st0 = snap()

        def feed_step(w, a):
            d.feed([w[0], w[1]] if stochastic else [w[1]])      # rock.py:443 gate first, then the sensor (404)
        r, t, acts, done, (x1, y1, st1) = run(env, d, e, T, 2, feed_step, snap)
        for key, v in zip(res, (x0, y0, st0, r, t, acts, done, x1, y1, st1)):
            res[key].append(v)
    for key in res:
        out[f"{tag}_{key}"] = pad(res[key], T) if key == "acts" else np.array(res[key])
    out[f"{tag}_cfg"] = np.array([n, k, int(stochastic), T])
    out[f"{tag}_discount"] = env._discount
    print(tag, "mean steps", np.mean(res["steps"]), "done", int(np.sum(res["done"])), "mean ret", np.mean(res["ret"]))


def gen_tag(E, out, tag, n_opp, M, T):
    d = ref_shim.draws()
    env = E.TagEnv(num_opponents=n_opp)
    g = env.grid
    res = {key: [] for key in ("agent0", "opp0", "ret", "steps", "acts", "done", "agent1", "opp1", "nopp1")}
    orig_move = env.move_opponent
    cur = {}

    def move_opponent(opp):                                      # slot j belongs to opponent j (tag.py:201-207)
        d.clear(); d.feed([cur["w"][opp], tag_pick_word(cur["w"][opp])])
        orig_move(opp)
        d.clear()
    env.move_opponent = move_opponent

    def snap():
        return g.get_index(env.state.agent_pos), [g.get_index(o) for o in env.state.opponent_pos], env.state.num_opp
    for e in range(M):
        rw = W(e, RESET_CTR, philox.DOMAIN_RESET, (1 + n_opp + 2) // 3)     # the j-th randint(29) = digit j % 3 of slot j // 3
        d.clear(); d.feed([tag_reset_word(int(rw[j // 3]), j % 3) for j in range(1 + n_opp)])
        env.reset()
        d.clear()
        a0, o0, _ = snap()

        def feed_step(w, a):
            cur["w"] = w
        r, t, acts, done, (a1, o1, n1) = run(env, d, e, T, n_opp, feed_step, snap, gamma=env._discount)
        for key, v in zip(res, (a0, o0, r, t, acts, done, a1, o1, n1)):
            res[key].append(v)
    env.move_opponent = orig_move
    for key in res:
        out[f"{tag}_{key}"] = pad(res[key], T) if key == "acts" else np.array(res[key])
    out[f"{tag}_cfg"] = np.array([n_opp, T])
    out[f"{tag}_discount"] = env._discount
    print(tag, "mean steps", np.mean(res["steps"]), "done", int(np.sum(res["done"])), "mean ret", np.mean(res["ret"]))


def gen_tiger(E, out, M, T):
    d = ref_shim.draws()
    env = E.TigerEnv()
    res = {key: [] for key in ("s0", "ret", "steps", "acts", "done", "s1")}
    for e in range(M):
        d.clear(); d.feed_gym(W(e, RESET_CTR, philox.DOMAIN_RESET, 1))
        env.reset()
        s0 = env.state

        def feed_step(w, a):
            d.feed_gym([w[0]]); d.feed([w[0]])                  # tiger.py:118-119 (gym's RNG), 143: ONE word serves both
        r, t, acts, done, s1 = run(env, d, e, T, 1, feed_step, lambda: env.state)
        for key, v in zip(res, (s0, r, t, acts, done, s1)):
            res[key].append(v)
    for key in res:
        out[f"tiger_{key}"] = pad(res[key], T) if key == "acts" else np.array(res[key])
    out["tiger_cfg"] = np.array([T])
    out["tiger_discount"] = env._discount
    print("tiger mean steps", np.mean(res["steps"]), "mean ret", np.mean(res["ret"]))


def gen_network(E, out, M, T):
    d = ref_shim.draws()
    n = 10
    env = E.NetworkEnv(n_machines=n, problem_type=3)
    res = {key: [] for key in ("ret", "steps", "acts", "s1")}
    for e in range(M):
        env.reset()

        def feed_step(w, a):                                    # one draw per UP machine in index order, then the action's
            d.feed([w[m] for m in range(n) if env.state[m]] + ([w[n]] if a < 2 * n else []))
        r, t, acts, done, s1 = run(env, d, e, T, n + 1, feed_step, lambda: int(sum(int(v) << m for m, v in enumerate(env.state))),
                                   step_words=lambda e_, ctr: philox.network_draws(SEED, np.array([e_]), ctr, n, env._p, env._q)[0])
        assert not done
        for key, v in zip(res, (r, t, acts, s1)):
            res[key].append(v)
    for key in res:
        out[f"network_{key}"] = pad(res[key], T) if key == "acts" else np.array(res[key])
    out["network_cfg"] = np.array([n, 3, T])
    out["network_discount"] = env._discount
    print("network mean ret", np.mean(res["ret"]))


def gen_battleship(E, out, tag, xs, ys, M, T):
    d = ref_shim.draws()
    env = E.BattleShipEnv(board_size=(xs, ys))
    res = {key: [] for key in ("occ", "ret", "steps", "acts", "done", "vis1", "rem1")}

    def board(attr):
        return np.array([[getattr(env.grid.board[x, y], attr) for y in range(ys)] for x in range(xs)])
    for e in range(M):
        d.clear(); d.feed(W(e, RESET_CTR, philox.DOMAIN_RESET, 600))    # rejection loop: attempt a -> slots 2a, 2a+1
        env.reset()
        d.clear()
        occ = board("occupied")
        r, t, acts, done, (vis1, rem1) = run(env, d, e, T, 1, lambda w, a: None,
                                             lambda: (board("visited"), env.state.total_remaining))
        for key, v in zip(res, (occ, r, t, acts, done, vis1, rem1)):
            res[key].append(v)
    for key in res:
        out[f"{tag}_{key}"] = pad(res[key], T) if key == "acts" else np.array(res[key])
    out[f"{tag}_cfg"] = np.array([xs, ys, 3, T])
    out[f"{tag}_discount"] = env._discount
    print(tag, "mean steps", np.mean(res["steps"]), "done", int(np.sum(res["done"])), "mean ret", np.mean(res["ret"]))


def main():
    E = ref_shim.load_reference()
    out = {"seed": SEED, "reset_ctr": RESET_CTR, "first_ctr": FIRST_CTR}
    with ref_shim.scripted_numpy():
        gen_rock(E, out, "rock_7_8", 7, 8, False, 96, 60)
        gen_rock(E, out, "rock_11_11", 11, 11, False, 96, 80)
        gen_rock(E, out, "rock_15_15", 15, 15, False, 64, 60)
        gen_rock(E, out, "srock_7_8", 7, 8, True, 64, 60)
        gen_tag(E, out, "tag_1opp", 1, 96, 90)
        gen_tag(E, out, "tag_2opp", 2, 48, 90)
        gen_tiger(E, out, 128, 30)
        gen_network(E, out, 48, 40)
        gen_battleship(E, out, "ship_5x5", 5, 5, 48, 30)
        gen_battleship(E, out, "ship_10x10", 10, 10, 48, 110)
    np.savez_compressed(os.path.join(GOLDEN, "rollouts.npz"), **out)
    print("wrote rollouts.npz")




def gen_rock_stats(E, out, tag, n, k, stochastic, M, T):
    """Belief side-statistics (rock.py:177-191) along check-heavy action sequences, recorded after every step."""
    import warnings
    warnings.simplefilter("ignore")
    d = ref_shim.draws()
    env = E.StochasticRockEnv(n, k) if stochastic else E.RockEnv(n, k)
    rs = np.random.RandomState(1000 + n + int(stochastic))
    acts = np.where(rs.rand(M, T) < 0.8, rs.randint(5, 5 + k, (M, T)), rs.randint(0, 4, (M, T)))
    acts[:, 5:] = np.where(rs.rand(M, T - 5) < 0.9, 5 + (np.arange(M)[:, None] % k), acts[:, 5:])
    acts[:, 5:][acts[:, 5:] < 5] = 5                # after five free steps only checks, 90 % of them hammering ONE rock:
                                                    # drives its lkv and lkw to underflow, where the reference's 0/0 gives NaN
    keys = ("count", "measured", "lkv", "lkw", "prob_valuable")
    rec = {key: np.zeros((M, T, k), np.float64) for key in keys}
    obs, alive = np.zeros((M, T), np.int32), np.zeros((M, T), bool)
    x0, y0, st0 = [], [], []
    for e in range(M):
        rw = W(e, RESET_CTR, philox.DOMAIN_RESET, 1)
        d.clear(); d.feed([rock_reset_word(int(rw[0]), r) for r in range(k)])
        env.reset()
        d.clear()
        x0.append(env.state.agent_pos.x); y0.append(env.state.agent_pos.y); st0.append([r.status for r in env.state.rocks])
        for t in range(T):
            w = W(e, FIRST_CTR + t, philox.DOMAIN_STEP, 2)
            d.clear(); d.feed([w[0], w[1]] if stochastic else [w[1]])
            ob, rw_, done, _ = env.step(int(acts[e, t]))
            d.clear()
            obs[e, t], alive[e, t] = ob, True
            for key in keys:
                rec[key][e, t] = [getattr(r, key) for r in env.state.rocks]
            if done:
                break
    out[f"{tag}_cfg"] = np.array([n, k, int(stochastic), T])
    out[f"{tag}_acts"], out[f"{tag}_obs"], out[f"{tag}_alive"] = acts.astype(np.int32), obs, alive
    out[f"{tag}_x0"], out[f"{tag}_y0"], out[f"{tag}_st0"] = np.array(x0), np.array(y0), np.array(st0)
    for key in keys:
        out[f"{tag}_{key}"] = rec[key]
    print(tag, "nan prob_valuable entries:", int(np.isnan(rec["prob_valuable"]).sum()), "min lkv", rec["lkv"][alive].min())


def gen_rock_heur(E, out, tag, n, k, stochastic, M, T, positional):
    """Heuristic rollouts (rock.py:557-572 with use_heuristic=True) played by the unmodified reference: its own
    History / Transition classes, ``np.random.choice(env._generate_preferred(history))`` scripted from the POLICY word.
    positional: transitions are built as the reference's own loop builds them, Transition(ob, action, next_ob, rw, done)
    (rock.py:566), which puts the reward into ``next_observation``; otherwise by field name.  Recorded per step: the
    preferred list (as a bit mask over action ids; the lists are in increasing action order) and whether it was the
    fallback to _generate_legal(), the action, the observation; per episode the return, steps and the final state."""
    import gym_pomdp.envs.rock as R
    d = ref_shim.draws()
    env = (E.StochasticRockEnv if stochastic else E.RockEnv)(n, k, use_heuristic=True)
    res = {key: [] for key in ("x0", "y0", "st0", "ret", "steps", "done", "x1", "y1", "st1")}
    masks, fallback, acts, obs = (-np.ones((M, T), np.int64) for _ in range(4))
    n_sample_branch = n_east_branch = n_fallback = 0
    for e in range(M):
        rw = W(e, RESET_CTR, philox.DOMAIN_RESET, 1)
        d.clear(); d.feed([rock_reset_word(int(rw[0]), r) for r in range(k)])
        ob = env.reset()
        d.clear()
        res["x0"].append(env.state.agent_pos.x); res["y0"].append(env.state.agent_pos.y)
        res["st0"].append([r.status for r in env.state.rocks])
        history = R.History()
        r, disc, t, done = 0.0, 1.0, 0, False
        while t < T and not done:
            ctr = FIRST_CTR + t
            calls = []
            orig_legal = env._generate_legal
            env._generate_legal = lambda: (calls.append(1), orig_legal())[1]     # with use_heuristic it is only reached by the
            pref = env._generate_preferred(history)                              # fallback `return self._generate_legal()`
            env._generate_legal = orig_legal
            is_fb = bool(calls)
            masks[e, t] = sum(1 << int(a) for a in set(pref))
            fallback[e, t] = int(is_fb)
            n_fallback += int(is_fb)
            n_sample_branch += int(list(pref) == [4])
            n_east_branch += int(list(pref) == [1])
            d.clear(); d.feed([W(e, ctr, philox.DOMAIN_POLICY, 1)[0]])
            a = int(np.random.choice(pref))
            d.clear()
            w = W(e, ctr, philox.DOMAIN_STEP, 2)
            d.feed([w[0], w[1]] if stochastic else [w[1]])
            next_ob, rw_, done, info = env.step(a)
            d.clear()
            if positional:
                history.append(R.Transition(ob, a, next_ob, rw_, done))           # rock.py:566 as written
            else:
                history.append(R.Transition(observation=ob, action=a, reward=rw_, next_observation=next_ob, done=done))
            ob = next_ob
            acts[e, t], obs[e, t] = a, next_ob
            r += rw_ * disc
            disc *= env._discount
            t += 1
        res["ret"].append(r); res["steps"].append(t); res["done"].append(bool(done))
        res["x1"].append(env.state.agent_pos.x); res["y1"].append(env.state.agent_pos.y)
        res["st1"].append([r_.status for r_ in env.state.rocks])
    for key in res:
        out[f"{tag}_{key}"] = np.array(res[key])
    out[f"{tag}_mask"], out[f"{tag}_fallback"], out[f"{tag}_acts"], out[f"{tag}_obs"] = masks, fallback, acts.astype(np.int32), obs.astype(np.int32)
    out[f"{tag}_cfg"] = np.array([n, k, int(stochastic), T, int(positional)])
    out[f"{tag}_discount"] = env._discount
    print(tag, "mean steps", np.mean(res["steps"]), "done", int(np.sum(res["done"])), "mean ret", np.mean(res["ret"]),
          "[SAMPLE]-only", n_sample_branch, "[EAST]-only", n_east_branch, "fallbacks", n_fallback)


class _TagHistory(object):
    """The History the reference's Tag loop expects (tag.py:303-313: ``history.append(action, ob)``, ``.size``,
    ``history[-1].ob`` / ``.action``); its module gym_pomdp.envs.history is not part of the repository."""

    class _E(object):
        def __init__(self, action, ob):
            self.action, self.ob = action, ob

    def __init__(self):
        self._h = []

    def append(self, action, ob):
        self._h.append(self._E(action, ob))

    def __getitem__(self, i):
        return self._h[i]

    @property
    def size(self):
        return len(self._h)


def gen_tag_heur(E, out, tag, n_opp, M, T):
    """tag.py:303-316 with ``env._generate_preferred(history)`` as the policy."""
    d = ref_shim.draws()
    env = E.TagEnv(num_opponents=n_opp)
    g = env.grid
    orig_move = env.move_opponent
    cur = {}

    def move_opponent(opp):
        d.clear(); d.feed([cur["w"][opp], tag_pick_word(cur["w"][opp])])
        orig_move(opp)
        d.clear()
    env.move_opponent = move_opponent
    res = {key: [] for key in ("agent0", "opp0", "ob0", "ret", "steps", "done", "agent1", "opp1", "nopp1")}
    masks, acts, obs = (-np.ones((M, T), np.int64) for _ in range(3))
    n_tag_only = 0
    for e in range(M):
        rw = W(e, RESET_CTR, philox.DOMAIN_RESET, (1 + n_opp + 2) // 3)
        d.clear(); d.feed([tag_reset_word(int(rw[j // 3]), j % 3) for j in range(1 + n_opp)])
        ob0 = env.reset()
        d.clear()
        res["agent0"].append(g.get_index(env.state.agent_pos)); res["opp0"].append([g.get_index(o) for o in env.state.opponent_pos])
        res["ob0"].append(ob0)
        history = _TagHistory()
        r, disc, t, done = 0.0, 1.0, 0, False
        while t < T and not done:
            ctr = FIRST_CTR + t
            pref = env._generate_preferred(history)
            masks[e, t] = sum(1 << int(a) for a in pref)
            n_tag_only += int(list(pref) == [4])
            d.clear(); d.feed([W(e, ctr, philox.DOMAIN_POLICY, 1)[0]])
            a = int(np.random.choice(pref))
            d.clear()
            cur["w"] = W(e, ctr, philox.DOMAIN_STEP, n_opp)
            ob, rw_, done, info = env.step(a)
            d.clear()
            history.append(a, ob)
            acts[e, t], obs[e, t] = a, ob
            r += rw_ * disc
            disc *= env._discount
            t += 1
        res["ret"].append(r); res["steps"].append(t); res["done"].append(bool(done))
        res["agent1"].append(g.get_index(env.state.agent_pos)); res["opp1"].append([g.get_index(o) for o in env.state.opponent_pos])
        res["nopp1"].append(env.state.num_opp)
    for key in res:
        out[f"{tag}_{key}"] = np.array(res[key])
    out[f"{tag}_mask"], out[f"{tag}_acts"], out[f"{tag}_obs"] = masks, acts.astype(np.int32), obs.astype(np.int32)
    out[f"{tag}_cfg"] = np.array([n_opp, T])
    out[f"{tag}_discount"] = env._discount
    print(tag, "mean steps", np.mean(res["steps"]), "done", int(np.sum(res["done"])), "mean ret", np.mean(res["ret"]), "[TAG]-only", n_tag_only)


def main_heur():
    E = ref_shim.load_reference()
    out = {"seed": SEED, "reset_ctr": RESET_CTR, "first_ctr": FIRST_CTR}
    with ref_shim.scripted_numpy():
        gen_rock_heur(E, out, "rock_7_8_fields", 7, 8, False, 96, 80, False)
        gen_rock_heur(E, out, "rock_7_8_main", 7, 8, False, 96, 80, True)
        gen_rock_heur(E, out, "rock_11_11_fields", 11, 11, False, 64, 100, False)
        gen_rock_heur(E, out, "rock_11_11_main", 11, 11, False, 64, 100, True)
        gen_rock_heur(E, out, "rock_4_3_fields", 4, 3, False, 64, 60, False)
        gen_rock_heur(E, out, "srock_7_8_fields", 7, 8, True, 64, 80, False)
        gen_tag_heur(E, out, "tag_1opp", 1, 96, 90)
        gen_tag_heur(E, out, "tag_2opp", 2, 48, 90)
    # RockEnv._select_target (rock.py:389-399) on random rock states of Rock(11,11)
    from gym_pomdp.envs.coord import Coord
    from gym_pomdp.envs.rock import config
    rs = np.random.RandomState(77)
    N = 400
    ax, ay = rs.randint(0, 11, N), rs.randint(0, 11, N)
    status, count = rs.randint(-1, 2, (N, 11)), rs.randint(-2, 3, (N, 11))
    tgt = []
    for i in range(N):
        st = type("S", (), {})()
        st.agent_pos = Coord(int(ax[i]), int(ay[i]))
        st.rocks = []
        for j in range(11):
            r = type("R", (), {})()
            r.status, r.count, r.pos = int(status[i, j]), int(count[i, j]), Coord(*config[11]["rock_pos"][j])
            st.rocks.append(r)
        tgt.append(E.RockEnv._select_target(st, 11))
    out.update(select_ax=ax, select_ay=ay, select_status=status, select_count=count, select_target=np.array(tgt))
    np.savez_compressed(os.path.join(GOLDEN, "heuristic_rollouts.npz"), **out)
    print("wrote heuristic_rollouts.npz")


def main_stats():
    E = ref_shim.load_reference()
    out = {"seed": SEED, "reset_ctr": RESET_CTR, "first_ctr": FIRST_CTR}
    with ref_shim.scripted_numpy():
        gen_rock_stats(E, out, "rock_7_8", 7, 8, False, 16, 3200)
        gen_rock_stats(E, out, "rock_15_15", 15, 15, False, 24, 200)
        gen_rock_stats(E, out, "srock_11_11", 11, 11, True, 24, 200)
    np.savez_compressed(os.path.join(GOLDEN, "rock_stats.npz"), **out)
    print("wrote rock_stats.npz")


if __name__ == "__main__":
    if "--heur-only" in sys.argv:
        main_heur()
    else:
        if "--stats-only" not in sys.argv:
            main()
            main_heur()
        main_stats()
