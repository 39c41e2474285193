"""CPU baseline leg of bench.py: the reference's own pure-Python RockSample step() loop (when its package is
reachable: /root/reference in the build container, oracle/_ref on the GPU box -- ``make -C oracle _ref``), else the
oracle's Python port of it, timed on host cores.

TEST/BENCH INFRASTRUCTURE ONLY -- imported by bench.py's ``cpu_baseline`` leg and by
``bench.py --impl reference``; never by gym_pomdp_b200/.

What is timed: for every (state, action) of a bounded sample of the benchmark's synthetic
workload, one call of ``oracle.pomdp_oracle.rock_step`` -- the scalar Python restatement of
rock.py:123-194, the same kind of code the reference runs (a Python ``step()`` per env
instance) minus its object/dict churn, so it is if anything FASTER than the reference's
own loop (reference 4.5e4 steps/s/core on Rock(11,11)).  Draw words come
from a pre-generated numpy array (cheaper than the reference's np.random.binomial call).
The unmodified reference itself cannot run on the GPU box (/root/reference is absent
there), hence ``kind = "port"``.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import pomdp_oracle as O


def rock_workload(n, k, count, seed=0x5EED):
    """The synthetic distribution of SURVEY.md §8d in the reference's own units."""
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    x = rs.randint(0, n, count)
    y = rs.randint(0, n, count)
    status = rs.randint(-1, 2, (count, k))
    action = rs.randint(0, 5 + k, count)
    words = rs.randint(0, 2 ** 32, (count, 2), dtype=np.uint64)
    return x, y, status, action, words


def _rock_loop(args):
    n, k, count, seed = args
    cfg = O.RockCfg(n, k)
    x, y, status, action, words = rock_workload(n, k, count, seed)
    xs, ys, acts = x.tolist(), y.tolist(), action.tolist()
    sts, ws = status.tolist(), words.tolist()
    step = O.rock_step
    acc = 0
    t0 = time.perf_counter()
    for i in range(count):
        w = ws[i]
        out = step(cfg, xs[i], ys[i], sts[i], acts[i], w.__getitem__)
        acc += out[4]
    dt = time.perf_counter() - t0
    return count, dt, acc


def time_rock(n, k, steps_per_proc, procs=None, seed=0x5EED):
    """Runs ``procs`` worker processes (default: every host core), each stepping its own
    ``steps_per_proc`` sampled envs once.  Returns dict(value=steps/s aggregate, ...)."""
    procs = procs or os.cpu_count() or 1
    jobs = [(n, k, steps_per_proc, seed + 7919 * p) for p in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_rock_loop(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_rock_loop, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)      # workers run concurrently: aggregate = total / slowest loop
    return {"value": total / slowest, "unit": "env-steps/s", "cores": procs, "kind": "port",
            "sample": "%d procs x %d RockSample(%d,%d) (state, action) pairs through oracle.pomdp_oracle.rock_step"
                      % (procs, steps_per_proc, n, k),
            "loop_seconds": slowest, "wall_seconds": wall, "steps": total}


# ---- persistent arm for ``bench.py --impl reference --steps K --warmup W`` -------------
_W = {}


def _arm_init(n, k, count, seed_base):
    ident = mp.current_process()._identity
    rank = ident[0] if ident else 0
    cfg = O.RockCfg(n, k)
    x, y, status, action, words = rock_workload(n, k, count, seed_base + 7919 * rank)
    _W.update(cfg=cfg, xs=x.tolist(), ys=y.tolist(), acts=action.tolist(), sts=status.tolist(), ws=words.tolist(),
              count=count)


def _arm_step(_):
    cfg, xs, ys, acts, sts, ws, count = (_W[k] for k in ("cfg", "xs", "ys", "acts", "sts", "ws", "count"))
    step = O.rock_step
    acc = 0
    t0 = time.perf_counter()
    for i in range(count):
        acc += step(cfg, xs[i], ys[i], sts[i], acts[i], ws[i].__getitem__)[4]
    return time.perf_counter() - t0, acc


class RockCpuArm(object):
    """Every host core steps its own fixed sample of ``count`` (state, action) pairs per step()."""

    def __init__(self, n, k, count, procs=None, seed=0x5EED):
        self.procs = procs or os.cpu_count() or 1
        self.count = count
        self.n, self.k = n, k
        self.pool = mp.get_context("fork").Pool(self.procs, initializer=_arm_init, initargs=(n, k, count, seed))

    def step(self):
        """One pass: returns (env-steps done, wall seconds of the pass)."""
        t0 = time.perf_counter()
        self.pool.map(_arm_step, range(self.procs), chunksize=1)
        return self.procs * self.count, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


# ---- the C restatement on every core: the strongest CPU form of the same algorithm ------------------
def _rock_c_loop(args):
    n, k, count, reps, seed = args
    from . import c_oracle as C
    x, y, status, action, words = rock_workload(n, k, count, seed)
    x, y, action = x.astype(np.int32), y.astype(np.int32), action.astype(np.int32)
    status, draws = status.astype(np.int8), words.astype(np.uint32)
    C.rock_step(n, k, False, 0.8, x[:64], y[:64], status[:64], action[:64], draws[:64])      # load the library
    t0 = time.perf_counter()
    acc = 0
    for _ in range(reps):
        out = C.rock_step(n, k, False, 0.8, x, y, status, action, draws)
        acc += int(out[3][0])
    return count * reps, time.perf_counter() - t0, acc


def time_rock_c(n, k, count=1 << 20, reps=8, procs=None, seed=0x5EED):
    """``oracle/pomdp_oracle.c`` (plain C, -O2, scalar) stepping ``count`` sampled envs ``reps`` times on every
    core, array copies of the ctypes wrapper included.  Reported beside the Python port as ``cpu_baseline.c_port``:
    the reference's own code is a Python loop, this is what the same algorithm does as compiled code.
    Returns None when the C oracle has not been built."""
    try:
        from . import c_oracle as C
        C.lib()
    except Exception:  # noqa: BLE001
        return None
    procs = procs or os.cpu_count() or 1
    jobs = [(n, k, count, reps, seed + 104729 * p) for p in range(procs)]
    if procs == 1:
        res = [_rock_c_loop(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_rock_c_loop, jobs)
    total, slowest = sum(r[0] for r in res), max(r[1] for r in res)
    return {"value": total / slowest, "unit": "env-steps/s", "cores": procs, "kind": "port",
            "sample": "%d procs x %d reps x %d RockSample(%d,%d) pairs through oracle/pomdp_oracle.c (oracle_rock_step)"
                      % (procs, reps, count, n, k)}


# ---- the UNMODIFIED reference: RockEnv._set_state + RockEnv.step (rock.py:243-245, 123-194) ----------------------
def _ref_states(E, n, k, count, seed):
    """The benchmark's synthetic (state, action) sample as the reference's own state dicts (rock.py:507-516)."""
    from gym_pomdp.envs.coord import Coord
    from gym_pomdp.envs.rock import config
    rock_pos = config[n]["rock_pos"]
    x, y, status, action, _ = rock_workload(n, k, count, seed)
    states = [{"agent_pos": (int(x[i]), int(y[i])), "target": -1,
               "rocks": [{"status": int(s), "pos": Coord(*rock_pos[j]), "count": 0, "measured": 0, "lkw": 1., "lkv": 1.,
                          "prob_valuable": .5} for j, s in enumerate(status[i])]} for i in range(count)]
    return states, action.tolist()


def _ref_pass(env, states, acts):
    """One pass of the planner pattern over the sample: env._set_state(s); env.step(a)."""
    set_state, step = env._set_state, env.step
    acc = 0
    t0 = time.perf_counter()
    for s, a in zip(states, acts):
        set_state(s)
        try:
            acc += step(a)[1]
        except IndexError:                       # Rock(15,15)'s dangling grid id at (12,2), rock.py:162
            pass
    return time.perf_counter() - t0, acc


def _ref_loop(args):
    n, k, count, seed = args
    from . import ref_shim
    E = ref_shim.load_reference()
    env = E.RockEnv(board_size=n, num_rocks=k)
    np.random.seed(seed & 0x7FFFFFFF)
    states, acts = _ref_states(E, n, k, count, seed)
    dt, acc = _ref_pass(env, states, acts)
    return count, dt, acc


def reference_available():
    from . import ref_shim
    return ref_shim.reference_available()


def time_rock_reference(n, k, steps_per_proc, procs=None, seed=0x5EED):
    """The reference's own RockEnv driven through ``_set_state`` + ``step`` on ``procs`` worker processes, each over
    its own sample of ``steps_per_proc`` (state, action) pairs, numpy's own RNG (unscripted).  kind = "reference"."""
    procs = procs or os.cpu_count() or 1
    jobs = [(n, k, steps_per_proc, seed + 7919 * p) for p in range(procs)]
    if procs == 1:
        res = [_ref_loop(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_ref_loop, jobs)
    total, slowest = sum(r[0] for r in res), max(r[1] for r in res)
    return {"value": total / slowest, "unit": "env-steps/s", "cores": procs, "kind": "reference",
            "sample": "%d procs x %d RockSample(%d,%d) (state, action) pairs through the unmodified reference: "
                      "RockEnv._set_state(s); RockEnv.step(a) (rock.py:243-245, 123-194), numpy RNG"
                      % (procs, steps_per_proc, n, k),
            "loop_seconds": slowest, "steps": total}


def _ref_arm_init(n, k, count, seed_base):
    from . import ref_shim
    ident = mp.current_process()._identity
    rank = ident[0] if ident else 0
    E = ref_shim.load_reference()
    np.random.seed((seed_base + 7919 * rank) & 0x7FFFFFFF)
    states, acts = _ref_states(E, n, k, count, seed_base + 7919 * rank)
    _W.update(env=E.RockEnv(board_size=n, num_rocks=k), states=states, acts=acts)


def _ref_arm_step(_):
    return _ref_pass(_W["env"], _W["states"], _W["acts"])


class RockReferenceArm(object):
    """``bench.py --impl reference``: every host core runs the unmodified reference over its own fixed sample of
    ``count`` (state, action) pairs per step()."""
    kind = "reference"

    def __init__(self, n, k, count, procs=None, seed=0x5EED):
        self.procs = procs or os.cpu_count() or 1
        self.count = count
        self.pool = mp.get_context("fork").Pool(self.procs, initializer=_ref_arm_init, initargs=(n, k, count, seed))

    def step(self):
        t0 = time.perf_counter()
        self.pool.map(_ref_arm_step, range(self.procs), chunksize=1)
        return self.procs * self.count, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
