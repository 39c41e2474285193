"""Shapes and corner cases of the batched entry points: empty and ragged batches, views that
are not 16-byte aligned, shard offsets that are not multiples of four, in-place steps,
masked resets and the per-element error flags that stand in for the reference's asserts
(rock.py:125-126, tag.py:109-110, battleship.py:93-95, tiger.py:74-75, network.py:73-74).

Runs on the host simulation here and, marked ``gpu``, on the CUDA library -- where the
same cases drive the vector path, the scalar path, the tail and the grid-stride loop of
pomdp_step_kernel / pomdp_reset_kernel against each other.
"""
import numpy as np
import pytest
import torch

import gym_pomdp_b200 as gp
from gym_pomdp_b200 import _lib

from backends import backend  # noqa: F401


def make_all(dev, B, seed=77, **kw):
    return {
        "rock": gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=B, device=dev, seed=seed, **kw),
        "rock15": gp.make("Rock-v0", board_size=15, num_rocks=15, batch_size=B, device=dev, seed=seed, **kw),
        "srock": gp.make("StochasticRock-v0", board_size=7, num_rocks=8, batch_size=B, device=dev, seed=seed, **kw),
        "tag": gp.make("Tag-v0", batch_size=B, device=dev, seed=seed, **kw),
        "tag3": gp.make("Tag-v0", num_opponents=3, batch_size=B, device=dev, seed=seed, **kw),
        "tiger": gp.make("Tiger-v0", batch_size=B, device=dev, seed=seed, **kw),
        "network": gp.make("Network-v0", batch_size=B, device=dev, seed=seed, **kw),
        "ship": gp.make("Battleship-v0", board_size=(10, 10), batch_size=B, device=dev, seed=seed, **kw),
    }


def random_inputs(env, name, n, rs, dev):
    """Synthetic (state, action) of SURVEY.md §8d for every env."""
    if name.startswith("rock") or name == "srock":
        b, k = env.grid.x_size, env.num_rocks
        state = env.pack(rs.randint(0, b, n), rs.randint(0, b, n), rs.randint(-1, 2, (n, k)))
        action = rs.randint(0, 5 + k, n)
    elif name.startswith("tag"):
        k = env.num_opponents
        state = env.pack(rs.randint(0, 29, n), rs.randint(0, 29, (n, k)))
        action = rs.randint(0, 5, n)
    elif name == "tiger":
        state, action = env.pack(rs.randint(0, 2, n)), rs.randint(0, 3, n)
    elif name == "network":
        state = torch.as_tensor(rs.randint(0, 1024, n), device=dev).int()
        action = rs.randint(0, 21, n)
    else:
        state, _ = env.init_states(n, step_ctr=99)
        action = rs.randint(0, 100, n)
    return state, torch.as_tensor(action, device=dev).int()


NAMES = ["rock", "rock15", "srock", "tag", "tag3", "tiger", "network", "ship"]


def eq(a, b):
    return all(torch.equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("name", NAMES)
def test_empty_batch(backend, name):
    env = make_all(backend, 0)[name]
    rs = np.random.RandomState(0)
    state, action = random_inputs(env, name, 0, rs, backend)
    out = env.simulate(state, action, step_ctr=1)
    assert [o.shape[0] for o in out] == [0, 0, 0, 0]
    st, ob = env.init_states(0)
    assert st.shape[0] == 0 and ob.shape[0] == 0
    assert int(env.belief_histogram(st).sum()) == 0


@pytest.mark.parametrize("name", NAMES)
def test_ragged_sizes_offsets_and_alignment_agree(backend, name):
    """One big aligned launch is the reference; every sub-range of it -- any length, any
    start (so global_offset % 4 != 0 and pointers off the 16-byte grid), i.e. the scalar
    kernel path and the 1..3-env tail -- must reproduce the same rows: the draw word of an
    env depends on its GLOBAL index only."""
    N = 1003
    env = make_all(backend, N)[name]
    rs = np.random.RandomState(5)
    state, action = random_inputs(env, name, N, rs, backend)
    full = env.simulate(state, action, step_ctr=9)
    for lo, hi in [(0, 1), (0, 3), (0, 4), (0, 5), (1, 2), (1, 9), (2, 1003), (3, 260), (4, 1001), (7, 7), (128, 640),
                   (998, 1003)]:
        env.global_offset = lo
        part = env.simulate(state[lo:hi], action[lo:hi], step_ctr=9)
        env.global_offset = 0
        assert eq(part, [f[lo:hi] for f in full]), (name, lo, hi)
    # reset: same property
    fs, fo = env.init_states(N, step_ctr=4)
    for lo, hi in [(0, 2), (1, 6), (5, 1003), (8, 1000)]:
        env.global_offset = lo
        ps = torch.empty_like(fs[lo:hi])
        po = torch.empty_like(fo[lo:hi])
        buf_s, buf_o = torch.empty_like(fs), torch.empty_like(fo)      # views into bigger buffers: odd alignment
        env.init_states(hi - lo, out=(buf_s[lo:hi], buf_o[lo:hi]), step_ctr=4)
        env.global_offset = 0
        assert torch.equal(buf_s[lo:hi], fs[lo:hi]) and torch.equal(buf_o[lo:hi], fo[lo:hi]), (name, lo, hi)
        del ps, po


@pytest.mark.parametrize("name", NAMES)
def test_in_place_step(backend, name):
    N = 517
    env = make_all(backend, N)[name]
    rs = np.random.RandomState(6)
    state, action = random_inputs(env, name, N, rs, backend)
    ref = env.simulate(state, action, step_ctr=3)
    work = state.clone()
    out = (work, torch.empty(N, dtype=torch.int32, device=backend), torch.empty(N, dtype=torch.float32, device=backend),
           torch.empty(N, dtype=torch.int32, device=backend))
    env.simulate(work, action, out=out, step_ctr=3)       # next_state aliases state
    assert eq(out, ref)


@pytest.mark.parametrize("name", NAMES)
def test_shard_invariance(backend, name):
    """Index-split shards (what N GPUs do) reproduce the single-device batch exactly."""
    N = 4096
    env = make_all(backend, N)[name]
    rs = np.random.RandomState(8)
    state, action = random_inputs(env, name, N, rs, backend)
    full = env.simulate(state, action, step_ctr=21)
    full_reset = env.init_states(N, step_ctr=22)
    for shards in (2, 8):
        per = N // shards
        parts, resets = [], []
        for r in range(shards):
            e = make_all(backend, per, global_offset=r * per)[name]
            parts.append(e.simulate(state[r * per:(r + 1) * per], action[r * per:(r + 1) * per], step_ctr=21))
            resets.append(e.init_states(per, step_ctr=22))
        cat = [torch.cat([p[i] for p in parts]) for i in range(4)]
        assert eq(cat, full), (name, shards)
        assert eq([torch.cat([p[i] for p in resets]) for i in range(2)], full_reset), (name, shards)


@pytest.mark.parametrize("name", NAMES)
def test_determinism_and_counter_separation(backend, name):
    N = 2048
    env = make_all(backend, N)[name]
    rs = np.random.RandomState(9)
    state, action = random_inputs(env, name, N, rs, backend)
    a = env.simulate(state, action, step_ctr=5)
    b = env.simulate(state, action, step_ctr=5)
    assert eq(a, b)
    if name not in ("ship",):                              # BattleShip's step draws nothing
        c = env.simulate(state, action, step_ctr=6)
        env2 = make_all(backend, N, seed=78)[name]
        d = env2.simulate(state, action, step_ctr=5)
        assert not eq(a, c) and not eq(a, d)
    r1, r2 = env.init_states(N, step_ctr=5), env.init_states(N, step_ctr=6)
    if name != "network":                                  # network.py:61-69: deterministic reset
        assert not torch.equal(r1[0], r2[0])


@pytest.mark.parametrize("name", NAMES)
def test_masked_reset(backend, name):
    N = 777
    env = make_all(backend, N)[name]
    env.reset()
    rs = np.random.RandomState(10)
    state, action = random_inputs(env, name, N, rs, backend)
    env._set_state(state)
    before = env.state.clone()
    mask = torch.as_tensor(rs.randint(0, 2, N), device=backend).bool()
    mask[:8] = torch.tensor([1, 0, 0, 0, 1, 1, 1, 1], device=backend).bool()    # a partial and a full group of four
    ctr = env._step_ctr + 1
    obs = env.reset(mask=mask)
    fresh, fresh_ob = env.init_states(N, step_ctr=ctr)
    m = mask.view(-1, *([1] * (before.dim() - 1)))
    assert torch.equal(env.state, torch.where(m, fresh, before))
    assert torch.equal(obs[mask], fresh_ob[mask])


def test_error_flags(backend):
    envs = make_all(backend, 8)
    dev = backend
    F = _lib
    # bad action / stepping a finished env / state outside the domain: flagged, state unchanged, obs = reward = 0
    for name in NAMES:
        env = envs[name]
        rs = np.random.RandomState(11)
        state, action = random_inputs(env, name, 8, rs, dev)
        bad = action.clone()
        bad[0], bad[3] = env.action_space.n, -1
        ns, ob, rw, fl = env.simulate(state, bad, step_ctr=2)
        for i in (0, 3):
            assert int(fl[i]) == F.FLAG_BAD_ACTION, (name, int(fl[i]))
            assert torch.equal(ns[i], state[i]) and int(ob[i]) == 0 and float(rw[i]) == 0.0
        good = env.simulate(state, action, step_ctr=2)
        for i in (1, 2, 4, 5, 6, 7):
            assert all(torch.equal(x[i], y[i]) for x, y in zip((ns, ob, rw, fl), good)), name
    # Rock: walk off the east edge -> done; stepping again -> STEPPED_DONE
    env = envs["rock"]
    s = env.pack([10] * 8, [3] * 8, np.ones((8, 11), int))
    ns, ob, rw, fl = env.simulate(s, torch.ones(8, dtype=torch.int32, device=dev), step_ctr=1)
    assert (rw == 10).all() and (fl == F.FLAG_DONE).all() and env.unpack(ns)[3].all()
    ns2, ob2, rw2, fl2 = env.simulate(ns, torch.zeros(8, dtype=torch.int32, device=dev), step_ctr=2)
    assert (fl2 == (F.FLAG_DONE | F.FLAG_STEPPED_DONE)).all() and torch.equal(ns2, ns) and (rw2 == 0).all()
    # Rock(15,15): the dangling grid id at (12, 2) (rock.py:60 lists 16 rocks, IndexError at rock.py:162)
    env = envs["rock15"]
    s = env.pack([12] * 8, [2] * 8, np.ones((8, 15), int))
    ns, ob, rw, fl = env.simulate(s, torch.full((8,), 4, dtype=torch.int32, device=dev), step_ctr=1)
    assert ((fl & F.FLAG_BAD_STATE) != 0).all() and (rw == -100).all()
    # Tag: agent cell 31 is not on the 29-cell board
    env = envs["tag"]
    s = env.pack([31] * 8, np.zeros((8, 1), int))
    ns, ob, rw, fl = env.simulate(s, torch.zeros(8, dtype=torch.int32, device=dev), step_ctr=1)
    assert (fl == F.FLAG_BAD_STATE).all() and torch.equal(ns, s)
    # Tiger: terminal step returns obs = state and a finished env
    env = envs["tiger"]
    s = env.pack([0, 1] * 4)
    a = torch.tensor([0, 1] * 4, dtype=torch.int32, device=dev)
    ns, ob, rw, fl = env.simulate(s, a, step_ctr=1)
    assert (rw == -20).all() and (fl == F.FLAG_DONE).all() and ob.tolist() == [0, 1] * 4       # tiger.py:81-83
    _, _, _, fl2 = env.simulate(ns, a, step_ctr=2)
    assert (fl2 == (F.FLAG_DONE | F.FLAG_STEPPED_DONE)).all()
    # Network: bits above n_machines are not a state
    env = envs["network"]
    s = torch.full((8,), 1 << 12, dtype=torch.int32, device=dev)
    ns, ob, rw, fl = env.simulate(s, torch.zeros(8, dtype=torch.int32, device=dev), step_ctr=1)
    assert (fl == F.FLAG_BAD_STATE).all() and torch.equal(ns, s)


def test_battleship_sink_everything(backend):
    """Shoot every cell of every board: reward -1 per fresh shot, +n_tiles on the last hit, done
    exactly when total_remaining hits 0, -10 on repeats, STEPPED_DONE afterwards."""
    B = 64
    env = gp.make("Battleship-v0", board_size=(10, 10), batch_size=B, device=backend, seed=3)
    env.reset()
    occ, _, _, _ = env.unpack(env.state)
    occ = occ.cpu().numpy().reshape(B, 10, 10)
    hits = np.zeros(B, int)
    finished = np.zeros(B, bool)
    for a in range(100):
        act = torch.full((B,), a, dtype=torch.int32, device=backend)
        ob, rw, done, info = env.step(act)
        ob, rw, done, fl = ob.cpu().numpy(), rw.cpu().numpy(), done.cpu().numpy(), info["flags"].cpu().numpy()
        live = ~finished
        hit = occ[:, a % 10, a // 10]
        hits[live] += hit[live]
        last = live & (hits == 5) & hit
        assert np.array_equal(ob[live], hit[live].astype(np.int32))
        assert np.array_equal(rw[live], np.where(last[live], 99.0, -1.0).astype(np.float32))
        assert np.array_equal(done[live], last[live])
        assert ((fl[finished] & _lib.FLAG_STEPPED_DONE) != 0).all()
        finished |= last
    assert finished.all()
    # repeat shot on an unfinished board costs -10 and reports a miss
    env.reset()
    a0 = torch.zeros(B, dtype=torch.int32, device=backend)
    env.step(a0)
    ob, rw, done, _ = env.step(a0)
    assert (rw == -10).all() and (ob == 0).all() and not done.any()


def test_tag_table_step_equals_the_general_functor_on_arbitrary_words(backend):
    """Stock Tag-v0: the one-table-word step (aligned batches: four envs per thread) against the general functor (the
    thread-per-env path an odd global offset selects) on ARBITRARY state words -- any num_opp field (0, negative in its
    6-bit two's complement, > 1), stray bits in the unused opponent fields, cell ids off the board, done states -- and
    actions 0..6.  Same global indices, hence the same draws; every output must agree."""
    n = 20000
    rs = np.random.RandomState(33)
    words = rs.randint(0, 1 << 32, n + 1, dtype=np.uint64)
    words[rs.rand(n + 1) < 0.5] &= 0x7FFFFFFF                     # half of them not done
    plausible = rs.rand(n + 1) < 0.6                               # most with cells on the board
    cells = rs.randint(0, 29, n + 1).astype(np.uint64) | (rs.randint(0, 29, n + 1).astype(np.uint64) << 5)
    words[plausible] = (words[plausible] & ~np.uint64(1023)) | cells[plausible]
    state = torch.from_numpy(words.astype(np.uint32).view(np.int32)).to(backend)
    action = torch.from_numpy(rs.randint(0, 7, n + 1).astype(np.int32)).to(backend)
    whole = gp.make("Tag-v0", batch_size=n + 1, device=backend, seed=77)
    shifted = gp.make("Tag-v0", batch_size=n, device=backend, seed=77, global_offset=1)
    a = whole.simulate(state, action, step_ctr=5)
    b = shifted.simulate(state[1:], action[1:], step_ctr=5)
    for x, y in zip(a, b):
        assert torch.equal(x[1:], y)
    fl = a[3].cpu().numpy()
    assert (fl & 2).any() and (fl & 4).any() and (fl & 8).any() and (fl == 0).any() and (fl == 1).any()   # every kind of flag occurred
    ap, bp = whole.simulate(state, action, step_ctr=5, packed=True), shifted.simulate(state[1:], action[1:], step_ctr=5, packed=True)
    assert torch.equal(ap[0][1:], bp[0]) and torch.equal(ap[1][1:], bp[1])


def test_belief_histogram_matches_bincount(backend):
    N = 5000
    envs = make_all(backend, N)
    rs = np.random.RandomState(12)
    for name in NAMES:
        env = envs[name]
        state, _ = random_inputs(env, name, N, rs, backend)
        h = env.belief_histogram(state).cpu().numpy()
        if name.startswith("rock") or name == "srock":
            x, y, st, _ = (v.cpu().numpy() for v in env.unpack(state))
            k = env.num_rocks
            exp = np.concatenate([(st == 1).sum(0), np.bincount(x | (y << 4), minlength=256)])
            assert np.array_equal(h, exp), name
        elif name.startswith("tag"):
            ag, op, _, _ = (v.cpu().numpy() for v in env.unpack(state))
            exp = np.concatenate([np.bincount(ag, minlength=29), np.bincount(op[:, 0], minlength=29)])
            assert np.array_equal(h, exp), name
        elif name == "tiger":
            assert np.array_equal(h, np.bincount(env.unpack(state)[0].cpu().numpy(), minlength=2))
        elif name == "network":
            s = state.cpu().numpy()
            assert np.array_equal(h, [((s >> m) & 1).sum() for m in range(10)])
        else:
            occ = env.unpack(state)[0].cpu().numpy().reshape(N, 10, 10)
            assert np.array_equal(h.reshape(10, 10), occ.sum(0).T)      # bin c = 10*y + x


def test_belief_histogram_once_is_self_cleaning(backend):
    """pomdp_belief_hist_once (what belief_histogram calls): one launch, the result OVERWRITES hist_out, the scratch is
    all zero again afterwards -- repeated calls over different particle sets and sizes (one CTA, many CTAs, ragged
    ends, an empty set onto a dirty output) equal the accumulate-into-zeroed-array entry point every time."""
    from gym_pomdp_b200 import _lib
    L = _lib.lib()
    envs = make_all(backend, 8)
    rs = np.random.RandomState(5)
    for name in NAMES:
        env = envs[name]
        p0, p1 = env._hist_args()
        bins = L.pomdp_belief_hist_bins(env.kind, p0, p1)
        scratch = torch.zeros(512 + 2, dtype=torch.int64, device=backend)
        out = torch.full((bins,), -7, dtype=torch.int64, device=backend)
        for n in (3000, 1, 0, 70001, 4096, 0, 5):
            state, _ = random_inputs(env, name, max(n, 1), rs, backend)
            state = state[:n]
            with env._guard():
                _lib.check(L.pomdp_belief_hist_once(env.kind, p0, p1, _lib.ptr(state), env.state_words, n, _lib.ptr(scratch),
                                                    _lib.ptr(out), env._stream()), "pomdp_belief_hist_once")
            exp = torch.zeros(bins, dtype=torch.int64, device=backend)
            with env._guard():
                _lib.check(L.pomdp_belief_hist(env.kind, p0, p1, _lib.ptr(state), env.state_words, n, _lib.ptr(exp), env._stream()),
                           "pomdp_belief_hist")
            assert torch.equal(out, exp), (name, n)
            assert not scratch.any(), (name, n)
            assert torch.equal(env.belief_histogram(state), exp), (name, n)
        assert L.pomdp_belief_hist_once(env.kind, p0, p1, _lib.ptr(state), env.state_words, 5, None, _lib.ptr(out), None) == -1
        assert L.pomdp_belief_hist_once(env.kind, p0, p1, _lib.ptr(state), env.state_words, 5, _lib.ptr(scratch), None, None) == -1


@pytest.mark.parametrize("name", NAMES)
def test_step_with_histogram_epilogue(backend, name):
    """simulate_hist (pomdp_E_step_hist: the step kernel counts the next states in its epilogue) returns exactly what
    simulate followed by belief_histogram(next_state) returns -- same draws, same results, same counts -- for aligned
    batches (four envs per thread), ragged ends, odd views (thread-per-env path), a single env, an empty batch, and call
    after call on the same self-cleaning scratch.  BattleShip runs its two kernels back to back behind the same call."""
    rs = np.random.RandomState(21)
    for n in (70001, 4096, 5, 1, 0, 33):
        env = make_all(backend, max(n, 1))[name]
        state, action = random_inputs(env, name, n + 1, rs, backend)
        for view in (slice(0, n), slice(1, n + 1)):              # the second view is not 16-byte aligned
            s, a = state[view], action[view]
            ref = env.simulate(s, a, step_ctr=9)
            ref_h = env.belief_histogram(ref[0])
            got = env.simulate_hist(s, a, step_ctr=9)
            assert eq(got[:4], ref), (name, n, view)
            assert torch.equal(got[4], ref_h), (name, n, view)
            assert int(got[4].sum()) == int(ref_h.sum())
        if hasattr(env, "_c_step_hist"):
            assert not env._local_hist_scratch().any()


def _fused_world_on_one_device(backend, env, state, world, calls, wait):
    """`world` ranks played on ONE device: each has its own symmetric buffer (slot 0 | slot 1 | arrivals) and scratch, all
    in one peer table.  wait=1 on CUDA: every rank's call goes to its own stream, so the kernels run concurrently and every
    kernel's last CTA waits for the arrivals of the others (hostsim counts arrivals but cannot wait: ranks run in turn)."""
    from gym_pomdp_b200 import _lib
    L = _lib.lib()
    p0, p1 = env._hist_args()
    bins = L.pomdp_belief_hist_bins(env.kind, p0, p1)
    n = state.shape[0]
    shard = [state[r * n // world:(r + 1) * n // world] for r in range(world)]
    bufs = [torch.zeros(2 * 512 + 64, dtype=torch.int64, device=backend) for _ in range(world)]
    table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=backend)
    scratch = [torch.zeros(514, dtype=torch.int64, device=backend) for _ in range(world)]
    outs = [torch.full((bins,), -1, dtype=torch.int64, device=backend) for _ in range(world)]
    cuda = torch.device(backend).type == "cuda"
    streams = [torch.cuda.Stream(backend) for _ in range(world)] if cuda else [None] * world
    if cuda:
        torch.cuda.synchronize()
    results = []
    for i in range(calls):
        used = []
        for r in range(world):
            sh = shard[r] if i % 3 != 2 or r != 1 else shard[r][:0]      # now and then a rank has an empty shard
            used.append(sh)
            h = streams[r].cuda_stream if cuda else None
            _lib.check(L.pomdp_belief_hist_allreduce(env.kind, p0, p1, _lib.ptr(sh), env.state_words, sh.shape[0],
                                                     _lib.ptr(scratch[r]), table.data_ptr(), world, r, wait, _lib.ptr(outs[r]), h),
                       "pomdp_belief_hist_allreduce")
        if cuda:
            torch.cuda.synchronize()
        slot = i & 1
        results.append({"slots": [b[slot * 512:(slot + 1) * 512].clone() for b in bufs], "other": [b[(slot ^ 1) * 512:(slot ^ 1) * 512 + 512].clone() for b in bufs],
                        "outs": [o.clone() for o in outs], "expected": torch.stack([env.belief_histogram(u) for u in used]).sum(0)})
    return results, bufs, scratch


@pytest.mark.parametrize("name", ["rock15", "tag", "network", "tiger", "ship"])
def test_fused_histogram_allreduce_adds_the_counts_into_every_peer_buffer(backend, name):
    """pomdp_belief_hist_allreduce without signalling (wait = 0: the caller would barrier): after every rank's call the
    current slot of EVERY rank's buffer holds the sum of the ranks' plain histograms, the other slot is cleared for the next
    call, the scratch comes back with zero counts and tickets and the call count, call after call."""
    N = 70001
    env = make_all(backend, N)[name]
    state, _ = random_inputs(env, name, N, np.random.RandomState(5), backend)
    world, calls = 3, 4
    results, bufs, scratch = _fused_world_on_one_device(backend, env, state, world, calls, wait=0)
    for i, res in enumerate(results):
        exp = res["expected"]
        for r in range(world):
            assert torch.equal(res["slots"][r][:exp.numel()], exp), (name, i, r)
            assert not res["slots"][r][exp.numel():].any() and not res["other"][r].any(), (name, i, r)
    for r in range(world):
        nb = exp.numel()                                                  # scratch: counts[bins] | tickets | calls made
        assert not scratch[r][:nb + 1].any() and int(scratch[r][nb + 1]) == calls and not scratch[r][nb + 2:].any()
        assert not bufs[r][1024:].any()                                   # nobody signalled


def test_fused_histogram_allreduce_with_in_kernel_signalling(backend):
    """wait = 1: three ranks on three streams.  Every call hands every rank the SAME global counts through hist_out
    (copied out by the kernel after it has seen all arrivals), and the arrival counters stand at the number of calls."""
    N = 3 * 4096 * 5 + 17
    envs = make_all(backend, N)
    rs = np.random.RandomState(6)
    for name in ("rock", "network"):
        env = envs[name]
        state, _ = random_inputs(env, name, N, rs, backend)
        world, calls = 3, 5
        results, bufs, _ = _fused_world_on_one_device(backend, env, state, world, calls, wait=1)
        for i, res in enumerate(results):
            exp = res["expected"]
            for r in range(world):
                assert torch.equal(res["slots"][r][:exp.numel()], exp), (name, i, r)
                if torch.device(backend).type == "cuda":                  # hostsim ranks run in turn: only the last one sees the total
                    assert torch.equal(res["outs"][r], exp), (name, i, r)
            assert torch.equal(res["outs"][world - 1], exp)
        for b in bufs:
            assert (b[1024:1024 + world] == calls).all() and not b[1024 + world:].any()


@pytest.mark.parametrize("size,max_len", [((10, 10), 3), ((5, 5), 3), ((10, 10), 5), ((12, 10), 4), ((4, 7), 3), ((3, 3), 3),
                                          ((120, 1), 4), ((1, 120), 4), ((2, 60), 5), ((30, 4), 6), ((40, 3), 9), ((8, 15), 9)])
def test_battleship_bitboard_reset_equals_warp_scan(backend, size, max_len):
    """The two fixed-time placement kernels (thread-per-env bitboards, warp-per-env ballot scan) enumerate the same
    accepted (pos, dir) set in the same order, so they must produce identical boards -- including 'no placement
    exists' (3x3 with ships 3 and 2), which both flag instead of spinning like the reference."""
    B = 3000
    a = gp.make("Battleship-v0", board_size=size, max_len=max_len, batch_size=B, device=backend, seed=9, reset_mode="scan")
    b = gp.make("Battleship-v0", board_size=size, max_len=max_len, batch_size=B, device=backend, seed=9, reset_mode="warpscan")
    c = gp.make("Battleship-v0", board_size=size, max_len=max_len, batch_size=B, device=backend, seed=9, reset_mode="table")
    sa, _ = a.init_states(B, step_ctr=4)
    sb, _ = b.init_states(B, step_ctr=4)
    assert torch.equal(sa, sb) and torch.equal(a.reset_flags, b.reset_flags)
    # the placement tables (accepted lists of ships 0 and 1 read instead of scanned; ships 2.. scanned): the same boards
    # through the TMA tile store, through per-thread stores (masked reset) and on an unaligned view
    sc, oc = c.init_states(B, step_ctr=4)
    assert torch.equal(sa, sc) and torch.equal(a.reset_flags, c.reset_flags) and not oc.any()
    mask = (torch.arange(B, device=sa.device) % 3 != 0).to(torch.uint8)
    sm = torch.full_like(sa, -1)
    c.init_states(B, out=(sm, torch.zeros(B, dtype=torch.int32, device=sa.device)), mask=mask, step_ctr=4)
    assert torch.equal(sm[mask.bool()], sa[mask.bool()]) and (sm[~mask.bool()] == -1).all()
    buf = torch.zeros(B * 8 + 1, dtype=torch.int32, device=sa.device)
    su = buf[1:].view(B, 8)
    c.init_states(B, out=(su, torch.zeros(B, dtype=torch.int32, device=sa.device)), step_ctr=4)
    assert torch.equal(su, sa)
    if size == (3, 3):
        assert (a.reset_flags == _lib.FLAG_BAD_STATE).all()
    else:
        placed = a.reset_flags == 0
        # a crowded board can run out of placements for the last ships (both kernels flag the same boards)
        assert placed.all() or max_len >= 9
        assert placed.float().mean() > 0.5
        occ, vis, rem, done = a.unpack(sa)
        n_cells = sum(range(2, max_len + 1))
        assert (occ.reshape(B, -1).sum(1)[placed] == n_cells).all() and (rem[placed] == n_cells).all()


@pytest.mark.parametrize("name", [n for n in NAMES if n != "ship"])
def test_packed_step_equals_unpacked(backend, name):
    """*_step_packed: same transition, obs | flags << 8 | reward_units << 16 in one stream -- on the vector path, the
    scalar path (odd offset) and with error flags present."""
    N = 1003
    env = make_all(backend, N)[name]
    rs = np.random.RandomState(13)
    state, action = random_inputs(env, name, N, rs, backend)
    action[5] = env.action_space.n                               # BAD_ACTION survives the packing
    ref = env.simulate(state, action, step_ctr=8)
    ns, res = env.simulate(state, action, step_ctr=8, packed=True)
    ob, rw, fl = env.unpack_result(res)
    assert torch.equal(ns, ref[0]) and torch.equal(ob, ref[1]) and torch.equal(rw, ref[2]) and torch.equal(fl, ref[3])
    env.global_offset = 3
    ref = env.simulate(state[3:], action[3:], step_ctr=8)
    ns, res = env.simulate(state[3:], action[3:], step_ctr=8, packed=True)
    env.global_offset = 0
    ob, rw, fl = env.unpack_result(res)
    assert torch.equal(ns, ref[0]) and torch.equal(ob, ref[1]) and torch.equal(rw, ref[2]) and torch.equal(fl, ref[3])
    # in place
    work = state.clone()
    env.simulate(work, action, out=(work, torch.empty(N, dtype=torch.int32, device=backend)), step_ctr=8, packed=True)
    assert torch.equal(work, env.simulate(state, action, step_ctr=8)[0])


def test_packed_reward_range(backend):
    """Extreme rewards fit the signed 16-bit unit field: Rock -100, Network 11 machines' worth minus a reboot."""
    env = make_all(backend, 8)["rock"]
    s = env.pack([0] * 8, [0] * 8, np.ones((8, 11), int))
    ns, res = env.simulate(s, torch.full((8,), 2, dtype=torch.int32, device=backend), step_ctr=1, packed=True)   # SOUTH off the board
    ob, rw, fl = env.unpack_result(res)
    assert (rw == -100).all() and (fl == 1).all() and (ob == 0).all()
    net = make_all(backend, 8)["network"]
    s = torch.full((8,), 1023, dtype=torch.int32, device=backend)
    a = torch.tensor([20, 0, 1, 3, 20, 20, 5, 7], dtype=torch.int32, device=backend)
    ref = net.simulate(s, a, step_ctr=1)
    ns, res = net.simulate(s, a, step_ctr=1, packed=True)
    ob, rw, fl = net.unpack_result(res)
    assert torch.equal(rw, ref[2]) and torch.equal(ob, ref[1])
    assert rw.tolist() == [np.float32(v) for v in (11.0, 10.9, 8.5, 8.5, 11.0, 11.0, 8.5, 8.5)]


def test_network_reward_conversion_is_correctly_rounded():
    """network_step_n turns the integer number of tenths into float32 with a multiply and two FMAs instead of a
    division; the result must be float32(t) / float32(10) (IEEE division, what the reference's double rounds to) for
    every value the 16-bit unit field can hold and far beyond."""
    import ctypes
    from backends import build_hostsim
    lib = ctypes.CDLL(build_hostsim())
    t = np.arange(-(1 << 21), (1 << 21) + 1, dtype=np.int32)
    out = np.empty(t.size, dtype=np.float32)
    lib.pomdp_hostsim_tenths_to_float(t.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(t.size))
    assert np.array_equal(out, t.astype(np.float32) / np.float32(10))
    assert np.array_equal(out, (t.astype(np.float64) / 10.0).astype(np.float32))     # float32(the reference's double)


def test_rock_reset_codes_one_lop3_equals_per_rock_definition():
    """Reset draws every rock's status from ONE word (rock i's uniform = rotl32(w, 30 - 2i) / 2^32): the kernel forms all
    sixteen 2-bit codes as 0x55555555 | (~w & 0xAAAAAAAA) plus a never-taken tie fix-up.  Checked against the per-rock
    definition sign(u - .5) on every single-bit word, their neighbours, and 2^22 random words; and against the
    Python oracle on the special words."""
    import ctypes
    from backends import build_hostsim
    from oracle import pomdp_oracle as O
    lib = ctypes.CDLL(build_hostsim())
    special = [0, 0xFFFFFFFF, 0xAAAAAAAA, 0x55555555]
    for b in range(32):
        special += [1 << b, (1 << b) - 1, ((1 << b) + 1) & 0xFFFFFFFF, (~(1 << b)) & 0xFFFFFFFF, (3 << b) & 0xFFFFFFFF]
    w = np.concatenate([np.array(special, dtype=np.uint32),
                        np.random.RandomState(3).randint(0, 1 << 32, 1 << 22, dtype=np.uint64).astype(np.uint32)])
    fast, slow = np.empty_like(w), np.empty_like(w)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.pomdp_hostsim_rock_reset_codes16(p(w), p(fast), p(slow), ctypes.c_int64(w.size))
    assert np.array_equal(fast, slow)
    # the vector kernel's four-env form (one shift + one LOP3 per env, ties handled out of line) on the same words
    from gym_pomdp_b200 import _lib as L_
    for board, k in [(7, 8), (11, 11), (15, 15)]:
        q = L_.RockParams(board, k, 0, 0, 0.8)
        ww = np.ascontiguousarray(w[: (w.size // 4) * 4])
        f4, s4 = np.empty(ww.size, np.uint64), np.empty(ww.size, np.uint64)
        assert lib.pomdp_hostsim_rock_reset4(ctypes.byref(q), p(ww), p(f4), p(s4), ctypes.c_int64(ww.size // 4)) == 0
        assert np.array_equal(f4, s4), (board, k)
    cfg = O.RockCfg(15, 15)
    for i in range(len(special)):
        _, _, status, _ = O.rock_reset(cfg, lambda s, i=i: int(w[i]))
        codes = [(0, 1, 0, 3)[st & 3] if st >= 0 else 3 for st in status]      # +1 -> 01, 0 -> 00, -1 -> 11
        assert [(int(fast[i]) >> (2 * r)) & 3 for r in range(15)] == codes, hex(int(w[i]))


@pytest.mark.parametrize("probs", [(0.0, 0.5, 1.0), (1.0, 1.0, 0.0), (0.25, 0.0, 0.5), (0.33, 0.1, 0.95), (1.0, 0.0, 0.5),
                                   (0.0, 0.0, 0.0), (0.5, 0.5, 0.5)])
def test_network_probabilities_of_exactly_zero_and_one(backend, probs):
    """Probabilities of exactly 0 and 1 (alias columns of weight 0, an outcome of weight 1), q < p (the larger
    probability then belongs to "no neighbour down") and p == q must agree with the oracle.  The reference hard-codes
    p, q, p_ob (network.py:27-38); the C ABI takes them as parameters."""
    from oracle import c_oracle as C, philox
    n, ptype, B = 10, 3, 4096
    env = gp.make("Network-v0", n_machines=n, problem_type=ptype, batch_size=B, device=backend, seed=11)
    env._params = _lib.NetworkParams(n, ptype, *probs)
    g = torch.Generator().manual_seed(5)
    s0 = torch.randint(0, 1 << n, (B,), generator=g).int()
    action = torch.randint(0, 2 * n + 1, (B,), generator=g).int()
    ns, ob, rw, fl = env.simulate(s0.to(backend), action.to(backend), step_ctr=3)
    bits = ((s0.numpy()[:, None] >> np.arange(n)) & 1).astype(np.int8)
    em, eob, erw = C.network_step(n, ptype, bits, action.numpy(), C.network_draws(11, 0, B, 3, n, probs[0], probs[1]), *probs)
    assert np.array_equal(ns.cpu().numpy(), (em.astype(np.int64) << np.arange(n)).sum(1).astype(np.int32))
    assert np.array_equal(ob.cpu().numpy(), eob)
    assert np.array_equal(rw.cpu().numpy(), erw.astype(np.float32))
    assert not fl.any()


@pytest.mark.parametrize("size,max_len", [((12, 10), 4), ((4, 7), 3), ((120, 1), 4), ((1, 120), 4), ((2, 60), 5), ((30, 4), 6),
                                          ((8, 15), 9), ((40, 3), 9), ((11, 10), 5), ((16, 7), 6)])
def test_battleship_bitboard_reset_equals_oracle_on_odd_boards(backend, size, max_len):
    """The two-word bitboard placement (run masks by doubling, both directions of an axis from one run mask, boards
    crossing the 64-bit boundary at every possible column) against the C oracle's cell-by-cell collision walk
    (battleship.py:195-211): same boards, same `no placement exists` flags, for shapes far from the stock 10 x 10."""
    from oracle import c_oracle as C, philox
    B = 2000
    eocc, erem, err = C.battleship_reset_scan(size[0], size[1], max_len, C.fill_env_draws(21, 0, B, 9, philox.DOMAIN_SHIP, max_len - 1))
    # all three fixed-time kernels: placement tables + TMA tile store, bitboard scan, warp scan (whose literal
    # seven-shift neighbour mask, ship_mark and ship_pack work on the same two-word boards -- no 128-bit integers
    # on the device: ships straddle bits 31/32, 63/64 and 95/96 of the board on these shapes)
    for mode in ("table", "scan", "warpscan"):
        env = gp.make("Battleship-v0", board_size=size, max_len=max_len, batch_size=B, device=backend, seed=21, reset_mode=mode)
        st, _ = env.init_states(B, step_ctr=9)
        occ, vis, rem, done = env.unpack(st)
        assert np.array_equal(occ.cpu().numpy(), eocc), mode
        assert np.array_equal(rem.cpu().numpy(), erem), mode
        assert np.array_equal(env.reset_flags.cpu().numpy() != 0, err != 0), mode
        assert not vis.any() and not done.any()
