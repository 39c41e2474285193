#!/usr/bin/env python
"""Can the step kernel stream straight from / to pinned host memory (zero-copy over PCIe)?  Times
pomdp_rock_step[_packed] called directly on pinned host tensors vs the chunked copy pipeline."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402

dev = torch.device("cuda", 0)
B = 1 << 22
env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=B, device=dev, seed=1)
g = torch.Generator(device=dev); g.manual_seed(0)
s = env.pack(torch.randint(0, 11, (B,), generator=g, device=dev), torch.randint(0, 11, (B,), generator=g, device=dev),
             torch.randint(-1, 2, (B, 11), generator=g, device=dev))
a = torch.randint(0, 16, (B,), generator=g, device=dev, dtype=torch.int32)
pin = dict(device="cpu", pin_memory=True)
hs, ha = s.cpu().pin_memory(), a.cpu().pin_memory()
ho = (torch.empty(B, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.float32, **pin),
      torch.empty(B, dtype=torch.int32, **pin))
hp = (torch.empty(B, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.int32, **pin))
ref = env.simulate(s, a, step_ctr=1)
res = {}


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


try:
    def zc_unpacked():
        env._c_step(hs, ha, ho[0], ho[1], ho[2], ho[3], B, 1)
    res["zero_copy_unpacked_ms"] = timeit(zc_unpacked)
    res["zero_copy_unpacked_ok"] = bool(torch.equal(ho[0], ref[0].cpu()) and torch.equal(ho[1], ref[1].cpu())
                                        and torch.equal(ho[2], ref[2].cpu()) and torch.equal(ho[3], ref[3].cpu()))

    def zc_packed():
        env._c_step_packed(hs, ha, hp[0], hp[1], B, 1)
    res["zero_copy_packed_ms"] = timeit(zc_packed)
    ob, rw, fl = env.unpack_result(hp[1])
    res["zero_copy_packed_ok"] = bool(torch.equal(hp[0], ref[0].cpu()) and torch.equal(ob, ref[1].cpu()) and torch.equal(rw, ref[2].cpu()))
    # device inputs, host outputs and vice versa
    d_ns, d_res = torch.empty_like(s), torch.empty(B, dtype=torch.int32, device=dev)
    res["host_in_dev_out_packed_ms"] = timeit(lambda: env._c_step_packed(hs, ha, d_ns, d_res, B, 1))
    res["dev_in_host_out_packed_ms"] = timeit(lambda: env._c_step_packed(s, a, hp[0], hp[1], B, 1))
except Exception as e:  # noqa: BLE001
    res["error"] = repr(e)
# (the hybrid modes -- copy engine one way, kernel loads / stores the other -- were measured slower than both pure
#  paths in round r01v and removed from simulate_host; "host_in_dev_out" / "dev_in_host_out" above time their kernels)
res["pipeline_packed_ms"] = timeit(lambda: env.simulate_host(hs, ha, hp, step_ctr=1, packed=True))                 # C host call
res["pipeline_packed_python_ms"] = timeit(lambda: env.simulate_host(hs, ha, hp, step_ctr=1, packed=True, pipeline="python"))
res["pipeline_unpacked_ms"] = timeit(lambda: env.simulate_host(hs, ha, ho, step_ctr=1))
for k in list(res):
    if k.endswith("_ms"):
        res[k.replace("_ms", "_steps_per_s")] = B / (res[k] * 1e-3)
print(json.dumps(res, indent=1))
