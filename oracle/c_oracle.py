"""numpy-facing ctypes wrapper of oracle/pomdp_oracle.c (oracle/_build/libpomdp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and bench.py's
cpu_baseline leg; never by gym_pomdp_b200/.  Build with ``make -C oracle``.
"""
import ctypes
import os
import subprocess
from ctypes import c_double, c_int, c_int64, c_uint32, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libpomdp_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "pomdp_oracle.c")
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = ctypes.CDLL(SO)
        _lib.oracle_rock_efficiency.restype = c_double
        _lib.oracle_rock_efficiency.argtypes = [c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def fill_draws(seed, global_offset, n, step, domain, n_slots):
    out = np.empty((n, n_slots), np.uint32)
    lib().oracle_fill_draws(c_uint64(seed), c_uint64(global_offset), c_int64(n), c_uint32(step), c_uint32(domain),
                            c_int(n_slots), _p(out))
    return out


def fill_env_draws(seed, global_offset, n, step, domain, n_slots):
    out = np.empty((n, n_slots), np.uint32)
    lib().oracle_fill_env_draws(c_uint64(seed), c_uint64(global_offset), c_int64(n), c_uint32(step), c_uint32(domain),
                            c_int(n_slots), _p(out))
    return out


def network_draws(seed, global_offset, n, step, n_machines, p=0.1, q=0.33):
    """uint32 [n, n_machines + 1]: the per-machine words + the observation word of Network's joint failure draw."""
    out = np.empty((n, n_machines + 1), np.uint32)
    lib().oracle_network_draws(c_uint64(seed), c_uint64(global_offset), c_int64(n), c_uint32(step), c_int(n_machines),
                               c_double(p), c_double(q), _p(out))
    return out


def network_alias(p=0.1, q=0.33):
    thr24, alias = np.empty(256, np.uint32), np.empty(256, np.int32)
    lib().oracle_network_alias(c_double(p), c_double(q), _p(thr24), _p(alias))
    return thr24, alias


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().oracle_philox_kat(_p(_c(ctr, np.uint32)), _p(_c(key, np.uint32)), _p(out))
    return out


def rock_grid(n, k):
    grid = np.empty((n, n), np.int8)
    pos = np.full((16, 2), -1, np.int32)
    start = np.zeros(2, np.int32)
    listed = lib().oracle_rock_grid(c_int(n), c_int(k), _p(grid), _p(pos), _p(start))
    assert listed >= 0, "unknown Rock configuration"
    return grid, pos[:listed], start


def rock_efficiency(d):
    return lib().oracle_rock_efficiency(int(d))


def rock_step(n, k, stochastic, p_move, x, y, status, action, draws):
    """Returns (x2, y2, status2, obs, reward, done, err); inputs are not modified."""
    N = len(action)
    x, y = _c(x, np.int32).copy(), _c(y, np.int32).copy()
    status = _c(status, np.int8).copy().reshape(N, k)
    obs, reward = np.empty(N, np.int32), np.empty(N, np.float64)
    done, err = np.empty(N, np.uint8), np.empty(N, np.uint8)
    rc = lib().oracle_rock_step(c_int(n), c_int(k), c_int(int(stochastic)), c_double(p_move), c_int64(N), _p(x), _p(y),
                                _p(status), _p(_c(action, np.int32)), _p(_c(draws, np.uint32)), _p(obs), _p(reward),
                                _p(done), _p(err))
    assert rc == 0
    return x, y, status, obs, reward, done.astype(bool), err


def rock_reset(n, k, draws):
    N = len(draws)
    x, y, obs = np.empty(N, np.int32), np.empty(N, np.int32), np.empty(N, np.int32)
    status = np.empty((N, k), np.int8)
    rc = lib().oracle_rock_reset(c_int(n), c_int(k), c_int64(N), _p(_c(draws, np.uint32)), _p(x), _p(y), _p(status), _p(obs))
    assert rc == 0
    return x, y, status, obs


def tag_admissible(agent, opp):
    out = np.empty(4, np.int8)
    lib().oracle_tag_admissible(c_int(agent), c_int(opp), _p(out))
    return out


def tag_step(n_opp, move_prob, agent, opp, num_opp, action, draws):
    N = len(action)
    agent = _c(agent, np.int32).copy()
    opp = _c(opp, np.int32).copy().reshape(N, n_opp)
    num_opp = _c(num_opp, np.int32).copy()
    obs, reward, done = np.empty(N, np.int32), np.empty(N, np.float64), np.empty(N, np.uint8)
    lib().oracle_tag_step(c_int(n_opp), c_double(move_prob), c_int64(N), _p(agent), _p(opp), _p(num_opp),
                          _p(_c(action, np.int32)), _p(_c(draws, np.uint32)), _p(obs), _p(reward), _p(done))
    return agent, opp, num_opp, obs, reward, done.astype(bool)


def tag_reset(n_opp, draws):
    N = len(draws)
    agent, opp = np.empty(N, np.int32), np.empty((N, n_opp), np.int32)
    num_opp, obs = np.empty(N, np.int32), np.empty(N, np.int32)
    lib().oracle_tag_reset(c_int(n_opp), c_int64(N), _p(_c(draws, np.uint32)), _p(agent), _p(opp), _p(num_opp), _p(obs))
    return agent, opp, num_opp, obs


def battleship_reset_rejection(xs, ys, max_len, draws):
    N, n_slots = draws.shape
    occ = np.empty((N, xs, ys), np.uint8)
    ships = np.zeros((N, max_len - 1, 4), np.int32)
    attempts, remaining = np.empty(N, np.int32), np.empty(N, np.int32)
    lib().oracle_battleship_reset_rejection(c_int(xs), c_int(ys), c_int(max_len), c_int64(N), _p(_c(draws, np.uint32)),
                                            c_int(n_slots), _p(occ), _p(ships), _p(attempts), _p(remaining))
    return occ.astype(bool), ships, attempts, remaining


def battleship_reset_scan(xs, ys, max_len, draws):
    N, n_slots = draws.shape
    occ = np.empty((N, xs, ys), np.uint8)
    remaining, err = np.empty(N, np.int32), np.empty(N, np.uint8)
    lib().oracle_battleship_reset_scan(c_int(xs), c_int(ys), c_int(max_len), c_int64(N), _p(_c(draws, np.uint32)),
                                       c_int(n_slots), _p(occ), _p(remaining), _p(err))
    return occ.astype(bool), remaining, err


def battleship_valid(xs, ys, occ, length):
    valid = np.empty(4 * xs * ys, np.uint8)
    cnt = lib().oracle_battleship_valid(c_int(xs), c_int(ys), _p(_c(occ, np.uint8)), c_int(length), _p(valid))
    return valid.astype(bool), cnt


def battleship_step(xs, ys, occ, vis, remaining, action):
    N = len(action)
    vis = _c(vis, np.uint8).copy().reshape(N, xs, ys)
    remaining = _c(remaining, np.int32).copy()
    obs, reward, done = np.empty(N, np.int32), np.empty(N, np.float64), np.empty(N, np.uint8)
    lib().oracle_battleship_step(c_int(xs), c_int(ys), c_int64(N), _p(_c(occ, np.uint8)), _p(vis), _p(remaining),
                                 _p(_c(action, np.int32)), _p(obs), _p(reward), _p(done))
    return vis.astype(bool), remaining, obs, reward, done.astype(bool)


def tiger_step(listen_prob, state, action, draws):
    N = len(action)
    state = _c(state, np.int32).copy()
    obs, reward, done = np.empty(N, np.int32), np.empty(N, np.float64), np.empty(N, np.uint8)
    lib().oracle_tiger_step(c_double(listen_prob), c_int64(N), _p(state), _p(_c(action, np.int32)),
                            _p(_c(draws, np.uint32)), _p(obs), _p(reward), _p(done))
    return state, obs, reward, done.astype(bool)


def tiger_reset(draws):
    N = len(draws)
    state, obs = np.empty(N, np.int32), np.empty(N, np.int32)
    lib().oracle_tiger_reset(c_int64(N), _p(_c(draws, np.uint32)), _p(state), _p(obs))
    return state, obs


def network_neighbours(n, problem_type):
    nb = np.empty((n, 3), np.int32)
    rc = lib().oracle_network_neighbours(c_int(n), c_int(problem_type), _p(nb))
    assert rc == 0
    return nb


def network_step(n, problem_type, machines, action, draws, p=0.1, q=0.33, p_ob=0.95):
    N = len(action)
    machines = _c(machines, np.int8).copy().reshape(N, n)
    obs, reward = np.empty(N, np.int32), np.empty(N, np.float64)
    rc = lib().oracle_network_step(c_int(n), c_int(problem_type), c_double(p), c_double(q), c_double(p_ob), c_int64(N),
                                   _p(machines), _p(_c(action, np.int32)), _p(_c(draws, np.uint32)), _p(obs), _p(reward))
    assert rc == 0
    return machines, obs, reward


# ---- legal actions + uniform-legal rollouts ------------------------
def rock_legal(n, k, x, y, status):
    legal = np.empty(40, np.int32)
    cnt = lib().oracle_rock_legal(c_int(n), c_int(k), c_int(int(x)), c_int(int(y)), _p(_c(status, np.int8)), _p(legal))
    assert cnt >= 0
    return legal[:cnt].tolist()


def _fa(first_action):
    return None if first_action is None else _p(_c(first_action, np.int32))


def rock_rollout(n, k, stochastic, p_move, x, y, status, seed, goff, ctr0, max_steps, gamma, first_action=None):
    """Returns (x, y, status, ret, steps, done, err) after the rollouts; inputs are not modified."""
    N = len(x)
    x, y = _c(x, np.int32).copy(), _c(y, np.int32).copy()
    status = _c(status, np.int8).copy().reshape(N, k)
    ret, steps = np.empty(N, np.float64), np.empty(N, np.int32)
    done, err = np.empty(N, np.uint8), np.empty(N, np.uint8)
    rc = lib().oracle_rock_rollout(c_int(n), c_int(k), c_int(int(stochastic)), c_double(p_move), c_int64(N), _p(x), _p(y),
                                   _p(status), c_uint64(seed), c_uint64(goff), c_uint32(ctr0), c_int(max_steps),
                                   c_double(gamma), _fa(first_action), _p(ret), _p(steps), _p(done), _p(err))
    assert rc == 0
    return x, y, status, ret, steps, done.astype(bool), err


def tag_rollout(n_opp, move_prob, agent, opp, num_opp, seed, goff, ctr0, max_steps, gamma, first_action=None):
    N = len(agent)
    agent = _c(agent, np.int32).copy()
    opp = _c(opp, np.int32).copy().reshape(N, n_opp)
    num_opp = _c(num_opp, np.int32).copy()
    ret, steps, done = np.empty(N, np.float64), np.empty(N, np.int32), np.empty(N, np.uint8)
    lib().oracle_tag_rollout(c_int(n_opp), c_double(move_prob), c_int64(N), _p(agent), _p(opp), _p(num_opp), c_uint64(seed),
                             c_uint64(goff), c_uint32(ctr0), c_int(max_steps), c_double(gamma), _fa(first_action), _p(ret), _p(steps), _p(done))
    return agent, opp, num_opp, ret, steps, done.astype(bool)


def tiger_rollout(listen_prob, state, seed, goff, ctr0, max_steps, gamma, first_action=None):
    N = len(state)
    state = _c(state, np.int32).copy()
    ret, steps, done = np.empty(N, np.float64), np.empty(N, np.int32), np.empty(N, np.uint8)
    lib().oracle_tiger_rollout(c_double(listen_prob), c_int64(N), _p(state), c_uint64(seed), c_uint64(goff), c_uint32(ctr0),
                               c_int(max_steps), c_double(gamma), _fa(first_action), _p(ret), _p(steps), _p(done))
    return state, ret, steps, done.astype(bool)


def network_rollout(n, problem_type, machines, seed, goff, ctr0, max_steps, gamma, p=0.1, q=0.33, p_ob=0.95, first_action=None):
    N = len(machines)
    machines = _c(machines, np.int8).copy().reshape(N, n)
    ret, steps = np.empty(N, np.float64), np.empty(N, np.int32)
    rc = lib().oracle_network_rollout(c_int(n), c_int(problem_type), c_double(p), c_double(q), c_double(p_ob), c_int64(N),
                                      _p(machines), c_uint64(seed), c_uint64(goff), c_uint32(ctr0), c_int(max_steps),
                                      c_double(gamma), _fa(first_action), _p(ret), _p(steps))
    assert rc == 0
    return machines, ret, steps


def battleship_rollout(xs, ys, occ, vis, remaining, seed, goff, ctr0, max_steps, gamma, first_action=None):
    N = len(remaining)
    vis = _c(vis, np.uint8).copy().reshape(N, xs, ys)
    remaining = _c(remaining, np.int32).copy()
    ret, steps, done = np.empty(N, np.float64), np.empty(N, np.int32), np.empty(N, np.uint8)
    lib().oracle_battleship_rollout(c_int(xs), c_int(ys), c_int64(N), _p(_c(occ, np.uint8)), _p(vis), _p(remaining),
                                    c_uint64(seed), c_uint64(goff), c_uint32(ctr0), c_int(max_steps), c_double(gamma),
                                    _fa(first_action), _p(ret), _p(steps), _p(done))
    return vis.astype(bool), remaining, ret, steps, done.astype(bool)
