"""The N > 1 path on CPU: two ranks over ``gloo`` (world_size 2, rendezvous on 127.0.0.1), each
holding one index shard of the global batch on the host build of the functors.

What the multi-GPU design promises (DESIGN.md "Multi-GPU") and this checks:
  * step/reset need NO collective: rank r simply uses global_offset = r * B/2, and the
    gathered shard results equal the single-process batch word for word;
  * the one collective on the path, the belief-histogram all-reduce, sums the per-shard
    int64 counts into the histogram of the whole particle set;
  * bench.py's timing reduction (MAX over ranks) and reference-arm rule (rank 0 only).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 6000          # per rank; not a multiple of the 512-thread CTA, is a multiple of 4
WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(name, env, n, seed):
    rs = np.random.RandomState(seed)
    if name == "rock":
        return env.pack(rs.randint(0, 11, n), rs.randint(0, 11, n), rs.randint(-1, 2, (n, 11))), rs.randint(0, 16, n)
    if name == "tag":
        return env.pack(rs.randint(0, 29, n), rs.randint(0, 29, (n, 1))), rs.randint(0, 5, n)
    if name == "network":
        return torch.as_tensor(rs.randint(0, 1024, n)).int(), rs.randint(0, 21, n)
    if name == "tiger":
        return env.pack(rs.randint(0, 2, n)), rs.randint(0, 3, n)
    raise KeyError(name)


def _make(name, n, goff):
    import gym_pomdp_b200 as gp
    ids = {"rock": ("Rock-v0", dict(board_size=11, num_rocks=11)), "tag": ("Tag-v0", {}), "network": ("Network-v0", {}),
           "tiger": ("Tiger-v0", {}), "ship": ("Battleship-v0", dict(board_size=(10, 10)))}
    env_id, kw = ids[name]
    return gp.make(env_id, batch_size=n, device="cpu", seed=0xABCD, global_offset=goff, **kw)


def _worker(rank, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    from gym_pomdp_b200 import _lib
    from backends import build_hostsim
    _lib._inject_for_tests(build_hostsim())
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        result = {}
        for name in ("rock", "tag", "network", "tiger"):
            whole = _make(name, WORLD * B, 0)
            state, action = _inputs(name, whole, WORLD * B, 5)          # every rank derives the same global inputs
            action = torch.as_tensor(action).int()
            lo, hi = rank * B, (rank + 1) * B
            env = _make(name, B, lo)                                    # this rank's shard
            ns, ob, rw, fl = env.simulate(state[lo:hi].contiguous(), action[lo:hi].contiguous(), step_ctr=4)
            st0, ob0 = env.init_states(B, step_ctr=5)
            # the host-buffer call on this rank's shard (chunks of 64 envs): same packed words as the device-side step
            hp = (torch.empty_like(ns), torch.empty_like(ob))
            env.simulate_host(state[lo:hi].contiguous(), action[lo:hi].contiguous(), hp, step_ctr=4, packed=True,
                              pipeline="c", chunk=64, n_streams=2)
            p_ob, p_rw, p_fl = env.unpack_result(hp[1])
            host_ok = torch.tensor([int(torch.equal(hp[0], ns) and torch.equal(p_ob, ob) and torch.equal(p_rw, rw)
                                        and torch.equal(p_fl, fl))])
            dist.all_reduce(host_ok, op=dist.ReduceOp.MIN)
            hist = env.belief_histogram(ns, all_reduce=True)            # the only collective
            # the same step with the histogram in its epilogue (pomdp_E_step_hist), counts summed over the ranks
            sh = env.simulate_hist(state[lo:hi].contiguous(), action[lo:hi].contiguous(), step_ctr=4, all_reduce=True)
            host_ok &= int(torch.equal(sh[0], ns) and torch.equal(sh[1], ob) and torch.equal(sh[4], hist))
            dist.all_reduce(host_ok, op=dist.ReduceOp.MIN)
            gathered = [torch.empty_like(ns) for _ in range(WORLD)]
            dist.all_gather(gathered, ns)
            g_ob = [torch.empty_like(ob) for _ in range(WORLD)]
            dist.all_gather(g_ob, ob)
            g_st0 = [torch.empty_like(st0) for _ in range(WORLD)]
            dist.all_gather(g_st0, st0)
            if rank == 0:
                w_ns, w_ob, w_rw, w_fl = whole.simulate(state, action, step_ctr=4)
                w_st0, _ = whole.init_states(WORLD * B, step_ctr=5)
                result[name] = bool(torch.equal(torch.cat(gathered), w_ns) and torch.equal(torch.cat(g_ob), w_ob)
                                    and torch.equal(torch.cat(g_st0), w_st0)
                                    and torch.equal(hist, whole.belief_histogram(w_ns)) and int(host_ok) == 1)
        # BattleShip: reset shards + histogram of occupied cells
        env = _make("ship", B, rank * B)
        st0, _ = env.init_states(B, step_ctr=6)
        hist = env.belief_histogram(st0, all_reduce=True)
        g = [torch.empty_like(st0) for _ in range(WORLD)]
        dist.all_gather(g, st0)
        if rank == 0:
            whole = _make("ship", WORLD * B, 0)
            w_st0, _ = whole.init_states(WORLD * B, step_ctr=6)
            result["ship"] = bool(torch.equal(torch.cat(g), w_st0) and torch.equal(hist, whole.belief_histogram(w_st0))
                                  and int(hist.sum()) == 5 * WORLD * B)
        # bench.py's reduction: the job time is the MAX over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            result["max_over_ranks"] = float(t.item()) == 10.0 + WORLD - 1
            torch.save(result, os.path.join(out_dir, "result.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_and_histogram_allreduce(tmp_path):
    mp.spawn(_worker, args=(_free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    res = torch.load(os.path.join(str(tmp_path), "result.pt"))
    assert res == {"rock": True, "tag": True, "network": True, "tiger": True, "ship": True, "max_over_ranks": True}, res
