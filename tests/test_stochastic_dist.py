"""Stochastic transitions: per-action outcome distributions of the kernels on N = 10^7 draws
(seed stated below) against

  (a) the reference's analytic probabilities (its own formulas: rock.py:383-387 efficiency,
      tag.py:201-207/260-280 move multiset, tiger.py:143-148, network.py:94-112), one-sample
      chi-square, and
  (b) outcome counts recorded from the UNMODIFIED reference running on numpy's own RNG
      (tests/golden/ref_dist.npz, written by oracle/gen_dist.py), two-sample chi-square.

Bar (BASELINE.json north_star: "match its per-action distribution to chi-sq p>0.01 on 10^7 draws"), applied twice:
  * pooled per action class: the statistics and degrees of freedom of the class's independent sub-tables add and the
    pooled p must exceed 0.01;
  * per sub-table (one (cell, action) table each): every single p must exceed 0.01 / n_subtables (Bonferroni), so one
    bad table cannot hide among a hundred good ones while the family-wise false-alarm rate stays at 1 %.
The minimum sub-table p of every pool is recorded with POMDP_DIST_REPORT.  The CPU suite runs the same cases at
N = 2*10^5 on the host build of the functors; the ``gpu`` run uses N = 10^7.
"""
import json
import os

import numpy as np
import pytest
import torch
from scipy import stats

import gym_pomdp_b200 as gp
from oracle import pomdp_oracle as O

from backends import backend  # noqa: F401

SEED = 0x5EED           # the stated seed
P_MIN = 0.01


@pytest.fixture
def N(backend):
    return 10_000_000 if backend.startswith("cuda") else 200_000


def report(name, backend, N, **pools):
    """With POMDP_DIST_REPORT=<file> set, append the pooled statistics (chi-square, dof, p) of a test and the p-value of
    every sub-table as a JSON line."""
    path = os.environ.get("POMDP_DIST_REPORT")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps({"test": name, "backend": backend, "draws_per_case": N, "seed": SEED,
                                "pools": {k: {"chi2": v.chi, "dof": v.dof, "p": v.p, "n_subtables": len(v.sub),
                                              "min_subtable_p": v.min_p, "bonferroni_bar": v.bar,
                                              "subtable_p": [round(q, 6) for q in v.sub]} for k, v in pools.items()}}) + "\n")


class Pool(object):
    """Chi-square over the cells of ONE action class.  Every ``one`` / ``two`` call adds one sub-table (one (cell, action)
    pair): its own p-value is kept, and its statistic and degrees of freedom are added to the pooled ones."""

    def __init__(self):
        self.chi, self.dof = 0.0, 0
        self.sub = []                       # p-value of every sub-table

    def _add(self, chi, dof):
        self.chi += chi
        self.dof += dof
        self.sub.append(float(stats.chi2.sf(chi, dof)))

    def one(self, counts, probs):
        """observed counts vs the reference's analytic probabilities"""
        counts, probs = np.asarray(counts, np.float64), np.asarray(probs, np.float64)
        keep = probs > 0
        assert counts[~keep].sum() == 0, "an outcome the reference cannot produce was observed: %s" % counts.tolist()
        if keep.sum() >= 2:
            e = probs[keep] / probs[keep].sum() * counts.sum()
            self._add(float(((counts[keep] - e) ** 2 / e).sum()), int(keep.sum()) - 1)
        return self

    def two(self, a, b):
        """observed counts vs counts recorded from the unmodified reference (homogeneity test)"""
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        keep = (a + b) > 0
        assert ((a > 0) == (b > 0))[(a + b) > 50].all(), "supports differ: %s vs %s" % (a.tolist(), b.tolist())
        if keep.sum() >= 2:
            self._add(float(stats.chi2_contingency(np.stack([a[keep], b[keep]]), correction=False)[0]), int(keep.sum()) - 1)
        return self

    @property
    def p(self):
        return float(stats.chi2.sf(self.chi, self.dof)) if self.dof else 1.0

    @property
    def min_p(self):
        return min(self.sub) if self.sub else 1.0

    @property
    def bar(self):
        """the per-sub-table bar: 0.01 shared out over the family (Bonferroni)"""
        return P_MIN / max(1, len(self.sub))

    def check(self, what=""):
        """both bars: pooled p > 0.01 and every sub-table's p > 0.01 / n_subtables"""
        assert self.p > P_MIN, (what, "pooled", self.p, self.chi, self.dof)
        assert self.min_p > self.bar, (what, "sub-table", self.min_p, self.bar, len(self.sub))
        return True


def one_sample_p(counts, probs):
    return Pool().one(counts, probs).p


def two_sample_p(a, b):
    return Pool().two(a, b).p


def bincount(t, n):
    return torch.bincount(t.reshape(-1).long(), minlength=n).cpu().numpy()


def full(n, v, dev):
    return torch.full((n,), int(v), dtype=torch.int32, device=dev)


def test_rock_sensor(golden, backend, N):
    """Check actions (rock.py:171-175, 383-387, 401-407) at L1 distances 0, 1, 4, 7, 12, 20 and on a
    collected rock."""
    ref = golden("ref_dist")
    env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=N, device=backend, seed=SEED)
    cfg = O.RockCfg(11, 11)
    analytic, vs_ref = Pool(), Pool()
    for c, (x, y, rock, status) in enumerate(ref["rock_cases"].tolist()):
        st = torch.ones((1, 11), dtype=torch.int64)
        st[0, rock] = status
        state = env.pack([x], [y], st).expand(N).contiguous()
        ns, ob, rw, fl = env.simulate(state, full(N, 5 + rock, backend), step_ctr=100 + c)
        assert torch.equal(ns, state) and not rw.any() and not fl.any()
        counts = bincount(ob, 3)
        d = O.l1_distance(x, y, *cfg.rock_pos[rock])
        eff = cfg.efficiency(d)
        probs = [0, 1 - eff, eff] if status == 1 else [0, eff, 1 - eff]     # status 0 reads BAD w.p. eff (rock.py:401-407)
        analytic.one(counts, probs)
        vs_ref.two(counts, ref["rock_obs_counts"][c])
        if d == 0:
            assert counts[2 if status == 1 else 1] == N              # eff(0) == 1.0 exactly
    report("rock_sensor", backend, N, analytic=analytic, vs_reference=vs_ref)
    analytic.check("rock sensor, analytic") and vs_ref.check("rock sensor, vs reference")


def test_rock_sensor_every_distance_and_action(backend, N):
    """All 11 check actions of RockSample(11,11) from uniformly random cells: per action, the
    chi-square over its (distance, reading) cells against eff(d)."""
    env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=N, device=backend, seed=SEED)
    cfg = O.RockCfg(11, 11)
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randint(0, 11, (N,), generator=g)
    y = torch.randint(0, 11, (N,), generator=g)
    status = torch.randint(0, 2, (N, 11), generator=g) * 2 - 1
    action = torch.randint(5, 16, (N,), generator=g)
    ns, ob, rw, fl = env.simulate(env.pack(x, y, status), action.to(backend).int(), step_ctr=7)
    ob = ob.cpu()
    rock = action - 5
    pos = torch.tensor(cfg.rock_pos[:11])
    d = (x - pos[rock, 0]).abs() + (y - pos[rock, 1]).abs()
    truth = torch.gather(status, 1, rock[:, None])[:, 0]
    correct = (ob == torch.where(truth == 1, 2, 1)).long()
    assert ((ob == 1) | (ob == 2)).all()
    for a in range(11):
        sel = rock == a
        tab = torch.bincount(d[sel] * 2 + correct[sel], minlength=42).numpy().reshape(21, 2)
        pool = Pool()
        for dd in range(21):
            eff = cfg.efficiency(dd)
            if tab[dd].sum():
                pool.one(tab[dd], [1 - eff, eff])
        report("rock_check_action_%d" % (5 + a), backend, N, analytic=pool)
        pool.check("check action %d" % (5 + a))


def test_rock_reset_and_stochastic_gate(golden, backend, N):
    ref = golden("ref_dist")
    env = gp.make("Rock-v0", board_size=7, num_rocks=8, batch_size=N, device=backend, seed=SEED)
    st, _ = env.init_states(N, step_ctr=3)
    _, _, status, _ = env.unpack(st)
    good = (status == 1).sum(0).cpu().numpy()
    assert ((status == 1) | (status == -1)).all()
    analytic, vs_ref = Pool(), Pool()
    R = int(ref["rock_reset_trials"])
    for i in range(8):                                                    # rock.py:78-80
        analytic.one([N - good[i], good[i]], [.5, .5])
        vs_ref.two([N - good[i], good[i]], [R - ref["rock_reset_good"][i], ref["rock_reset_good"][i]])
    report("rock_reset_status", backend, N, analytic=analytic, vs_reference=vs_ref)
    analytic.check("rock reset, analytic") and vs_ref.check("rock reset, vs reference")
    env = gp.make("StochasticRock-v0", board_size=7, num_rocks=8, batch_size=N, device=backend, seed=SEED)
    state = env.pack([3], [3], torch.ones((1, 8), dtype=torch.int64)).expand(N).contiguous()
    ns, ob, rw, fl = env.simulate(state, full(N, 0, backend), step_ctr=4)
    moved = int((env.unpack(ns)[1] == 4).sum())
    gate_a, gate_r = Pool().one([N - moved, moved], [.2, .8]), Pool().two([N - moved, moved], ref["srock_moved"])
    report("stochastic_rock_p_move_gate", backend, N, analytic=gate_a, vs_reference=gate_r)
    gate_a.check("p_move gate") and gate_r.check("p_move gate vs reference")      # rock.py:429, 443


def tag_move_probs(a, o):
    """tag.py:201-207, 260-280: stay w.p. 0.2; else uniform over the multiset; a move off the board stays put"""
    ax, ay = O.tag_get_coord(a)
    ox, oy = O.tag_get_coord(o)
    acts = O.tag_admissible(ax, ay, ox, oy)
    probs = np.zeros(29)
    probs[o] += 0.2
    for m in acts:
        dx, dy = O.MOVES[m]
        probs[O.tag_get_index(ox + dx, oy + dy) if O.tag_is_inside(ox + dx, oy + dy) else o] += 0.8 / len(acts)
    return probs


def test_tag_opponent_move(golden, backend, N):
    ref = golden("ref_dist")
    env = gp.make("Tag-v0", batch_size=N, device=backend, seed=SEED)
    analytic, vs_ref = Pool(), Pool()
    for c, (a, o) in enumerate(ref["tag_cases"].tolist()):
        state = env.pack([a], [[o]]).expand(N).contiguous()
        ns, ob, rw, fl = env.simulate(state, full(N, 4, backend), step_ctr=200 + c)
        assert (rw == -10).all() and not fl.any() and (ob == a).all()
        counts = bincount(env.unpack(ns)[1][:, 0], 29)
        analytic.one(counts, tag_move_probs(a, o))
        vs_ref.two(counts, ref["tag_opp_counts"][c])
    report("tag_opponent_move_probes", backend, N, analytic=analytic, vs_reference=vs_ref)
    analytic.check("tag move probes") and vs_ref.check("tag move probes vs reference")


def test_tag_all_pairs(backend, N):
    """Failed TAG from every one of the 29*28 (agent != opponent) pairs: pooled chi-square."""
    env = gp.make("Tag-v0", batch_size=N, device=backend, seed=SEED)
    g = torch.Generator(device="cpu").manual_seed(2)
    a = torch.randint(0, 29, (N,), generator=g)
    o = (a + torch.randint(1, 29, (N,), generator=g)) % 29
    ns, ob, rw, fl = env.simulate(env.pack(a, o[:, None]), full(N, 4, backend), step_ctr=9)
    opp2 = env.unpack(ns)[1][:, 0].cpu()
    tab = torch.bincount((a * 29 + o) * 29 + opp2, minlength=29 ** 3).numpy().reshape(29, 29, 29)
    pool = Pool()
    for ai in range(29):
        for oi in range(29):
            if ai != oi:
                pool.one(tab[ai, oi], tag_move_probs(ai, oi))
    report("tag_opponent_move_all_812_pairs", backend, N, analytic=pool)
    pool.check("tag, all 812 pairs")


def test_tag_reset(golden, backend, N):
    ref = golden("ref_dist")
    env = gp.make("Tag-v0", batch_size=N, device=backend, seed=SEED)
    st, ob = env.init_states(N, step_ctr=15)
    agent, opp, nop, done = env.unpack(st)
    ca, co, cob = bincount(agent, 29), bincount(opp[:, 0], 29), bincount(ob, 30)
    analytic = Pool().one(ca, np.full(29, 1 / 29)).one(co, np.full(29, 1 / 29))      # tag.py:43-44, 181-193
    vs_ref = Pool().two(ca, ref["tag_reset_agent"]).two(co, ref["tag_reset_opp"])
    report("tag_reset_cells", backend, N, analytic=analytic, vs_reference=vs_ref)
    analytic.check("tag reset") and vs_ref.check("tag reset vs reference")
    # the reset observation is 29 exactly when agent and opponent coincide (tag.py:101, 219-226)
    assert torch.equal(ob == 29, agent == opp[:, 0]) and torch.equal(ob[ob != 29], agent[ob != 29])
    assert two_sample_p([N - cob[29], cob[29]], [ref["tag_reset_obs"][:29].sum(), ref["tag_reset_obs"][29]]) > P_MIN


def test_tiger(golden, backend, N):
    ref = golden("ref_dist")
    env = gp.make("Tiger-v0", batch_size=N, device=backend, seed=SEED)
    listen_a, listen_r, samp_a, samp_r = Pool(), Pool(), Pool(), Pool()
    for s in (0, 1):
        state = env.pack([s]).expand(N).contiguous()
        ns, ob, rw, fl = env.simulate(state, full(N, 2, backend), step_ctr=10 + s)        # listen
        assert torch.equal(ns, state) and (rw == -1).all()
        counts = bincount(ob, 3)
        listen_a.one(counts, [.85, .15, 0] if s == 0 else [.15, .85, 0])                  # tiger.py:141-148
        listen_r.two(counts, ref["tiger_listen"][s])
        ns, ob, rw, fl = env.simulate(state, full(N, 1 - s, backend), step_ctr=12 + s)    # open the safe door
        assert (ob == 2).all() and (rw == 10).all() and not fl.any()
        counts = bincount(env.unpack(ns)[0], 2)
        samp_a.one(counts, [.5, .5])                                                     # tiger.py:117-119
        samp_r.two(counts, ref["tiger_resample"][s])
    st, ob = env.init_states(N, step_ctr=14)
    counts = bincount(env.unpack(st)[0], 2)
    samp_a.one(counts, [.5, .5])
    samp_r.two(counts, ref["tiger_reset"])
    report("tiger", backend, N, listen_analytic=listen_a, listen_vs_reference=listen_r, resample_analytic=samp_a,
           resample_vs_reference=samp_r)
    for q, what in ((listen_a, "listen"), (listen_r, "listen vs reference"), (samp_a, "resample"), (samp_r, "resample vs reference")):
        q.check("tiger " + what)


def test_network(golden, backend, N):
    ref = golden("ref_dist")
    env = gp.make("Network-v0", batch_size=N, device=backend, seed=SEED)
    nb = O.network_neighbours(10, 3)
    T = int(ref["trials"])
    fail_a, fail_r = Pool(), Pool()
    for c, s in enumerate(ref["network_states"].tolist()):
        ns, ob, rw, fl = env.simulate(full(N, s, backend), full(N, 20, backend), step_ctr=20 + c)
        assert (ob == 2).all()
        for m in range(10):
            up = int(((ns >> m) & 1).sum())
            if not (s >> m) & 1:
                assert up == 0                                                      # down machines stay down
                continue
            p_fail = .33 if any(not (s >> j) & 1 for j in nb[m]) else .1             # network.py:94-99
            fail_a.one([N - up, up], [p_fail, 1 - p_fail])
            fail_r.two([N - up, up], [T - ref["network_up"][c, m], ref["network_up"][c, m]])
    report("network_failures", backend, N, analytic=fail_a, vs_reference=fail_r)
    fail_a.check("network failures") and fail_r.check("network failures vs reference")
    allup = full(N, 1023, backend)
    ns, ob, rw, fl = env.simulate(allup, full(N, 2, backend), step_ctr=30)           # ping machine 1
    bit = (ns >> 1) & 1
    tab = bincount(bit * 3 + ob, 6).reshape(2, 3)
    ob_a, ob_r = Pool(), Pool()
    for b in (0, 1):
        ob_a.one(tab[b], [.95, .05, 0] if b == 0 else [.05, .95, 0])                 # network.py:106-112
        ob_r.two(tab[b], ref["network_ping"][b])
    ns, ob, rw, fl = env.simulate(allup, full(N, 3, backend), step_ctr=31)           # reboot machine 1
    assert (((ns >> 1) & 1) == 1).all()
    counts = bincount(ob, 3)
    ob_a.one(counts, [.05, .95, 0])                                                 # network.py:101-105
    ob_r.two(counts, ref["network_reboot"])
    report("network_ping_reboot_obs", backend, N, analytic=ob_a, vs_reference=ob_r)
    ob_a.check("network observations") and ob_r.check("network observations vs reference")


@pytest.mark.parametrize("size", [(5, 5), (10, 10)])
def test_battleship_first_ship_placement(golden, backend, N, size):
    """The warp-scan reset against the reference's rejection loop: the set of cells covered by
    the length-3 ship (placements (pos, N) and (pos + 2N, S) cover the same cells)."""
    ref = golden("ref_dist")
    xs, ys = size
    B = min(N, 1 << 21)
    env = gp.make("Battleship-v0", board_size=size, batch_size=B, device=backend, seed=SEED)
    st, _ = env.init_states(B, step_ctr=6)
    occ = env.unpack(st)[0].reshape(B, xs, ys)
    assert (occ.reshape(B, -1).sum(1) == 5).all()
    horiz = occ[:, :-2, :] & occ[:, 1:-1, :] & occ[:, 2:, :]          # run of three along x starting at (x, y)
    vert = occ[:, :, :-2] & occ[:, :, 1:-1] & occ[:, :, 2:]
    assert ((horiz.reshape(B, -1).sum(1) + vert.reshape(B, -1).sum(1)) == 1).all()
    hc = horiz.sum(0).cpu().numpy()                                    # [xs-2, ys]
    vc = vert.sum(0).cpu().numpy()                                     # [xs, ys-2]
    first = ref[f"ship_{xs}x{ys}_first"]
    rh, rv = np.zeros_like(hc), np.zeros_like(vc)
    mh, mv = np.zeros_like(hc), np.zeros_like(vc)                      # how many valid placements cover the cell set
    for cand, cnt in enumerate(first.tolist()):
        pos, d = cand >> 2, cand & 3
        x, y = pos % xs, pos // xs
        if cnt == 0:
            continue
        tgt, mul, ix = ((rv, mv, (x, y)), (rh, mh, (x, y)), (rv, mv, (x, y - 2)), (rh, mh, (x - 2, y)))[d]
        tgt[ix] += cnt
        mul[ix] += 1
    kc, rc = np.concatenate([hc.ravel(), vc.ravel()]), np.concatenate([rh.ravel(), rv.ravel()])
    assert ((kc > 0) == (rc > 0)).all()                                # exactly the reference's support
    ship_r = Pool().two(kc, rc)
    ship_r.check("first ship vs reference")
    # uniform over the valid (pos, dir) set == each reachable cell set weighted by how many placements cover it
    assert int((first > 0).sum()) == (20 if size == (5, 5) else 240)
    mult = np.concatenate([mh.ravel(), mv.ravel()]).astype(np.float64)
    ship_a = Pool().one(kc, mult / mult.sum())
    report("battleship_first_ship_%dx%d" % size, backend, B, analytic=ship_a, vs_reference=ship_r)
    ship_a.check("first ship, analytic")
