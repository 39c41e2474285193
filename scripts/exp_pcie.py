#!/usr/bin/env python
"""PCIe characterisation for the e2e path: pinned H2D / D2H bandwidth alone and together, and simulate_host with
different chunk sizes.  python scripts/exp_pcie.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402

dev = torch.device("cuda", 0)
MB = 1 << 20


def bw(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


out = {}
for size in (4 * MB, 16 * MB, 64 * MB, 256 * MB):
    h = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(size, dtype=torch.uint8, device=dev)
    h2 = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(size, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out["h2d_%dMB" % (size // MB)] = bw(lambda: d.copy_(h, non_blocking=True), size)
    out["d2h_%dMB" % (size // MB)] = bw(lambda: h.copy_(d, non_blocking=True), size)

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    out["both_%dMB_total" % (size // MB)] = bw(both, 2 * size)
print(json.dumps(out))

B = 1 << 22
env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=B, device=dev, seed=1)
g = torch.Generator(device=dev); g.manual_seed(0)
s = env.pack(torch.randint(0, 11, (B,), generator=g, device=dev), torch.randint(0, 11, (B,), generator=g, device=dev),
             torch.randint(-1, 2, (B, 11), generator=g, device=dev))
a = torch.randint(0, 16, (B,), generator=g, device=dev, dtype=torch.int32)
hs, ha = s.cpu().pin_memory(), a.cpu().pin_memory()
pin = dict(device="cpu", pin_memory=True)
ho = (torch.empty(B, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.float32, **pin),
      torch.empty(B, dtype=torch.int32, **pin))
hp = (torch.empty(B, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.int32, **pin))
res = {}
for packed in (False, True):
    for ns in (2, 3, 4):
        for chunk in (1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21):
            o = hp if packed else ho
            for _ in range(3):
                env.simulate_host(hs, ha, o, step_ctr=1, chunk=chunk, packed=packed, n_streams=ns, pipeline="python")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                env.simulate_host(hs, ha, o, step_ctr=1, chunk=chunk, packed=packed, n_streams=ns, pipeline="python")
            dt = (time.perf_counter() - t0) / 20
            res["packed=%d streams=%d chunk=2^%d" % (packed, ns, chunk.bit_length() - 1)] = round(dt * 1e3, 3)
print(json.dumps(res, indent=0))
