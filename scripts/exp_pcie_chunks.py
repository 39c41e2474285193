#!/usr/bin/env python
"""Where the per-chunk cost of the host pipeline comes from: the copy pattern of simulate_host WITHOUT kernels
(two H2D + two D2H copies per chunk, slot-per-stream), one direction at a time and both, for several chunk sizes.
python scripts/exp_pcie_chunks.py"""
import json
import time

import torch

dev = torch.device("cuda", 0)
B = 1 << 22
pin = dict(device="cpu", pin_memory=True)
h_in = [torch.empty(B, dtype=torch.int32, **pin) for _ in range(2)]
h_out = [torch.empty(B, dtype=torch.int32, **pin) for _ in range(2)]
res = {}
for lg in (16, 18, 19, 20, 21):
    C = 1 << lg
    for ns in (2, 3):
        streams = [torch.cuda.Stream(dev) for _ in range(ns)]
        slots = [[torch.empty(C, dtype=torch.int32, device=dev) for _ in range(4)] for _ in range(ns)]
        for mode in ("h2d", "d2h", "both"):
            def run():
                for ci, lo in enumerate(range(0, B, C)):
                    k = ci % ns
                    with torch.cuda.stream(streams[k]):
                        d = slots[k]
                        if mode != "d2h":
                            d[0].copy_(h_in[0][lo:lo + C], non_blocking=True)
                            d[1].copy_(h_in[1][lo:lo + C], non_blocking=True)
                        if mode != "h2d":
                            h_out[0][lo:lo + C].copy_(d[2], non_blocking=True)
                            h_out[1][lo:lo + C].copy_(d[3], non_blocking=True)
                torch.cuda.synchronize()
            run(); run()
            t0 = time.perf_counter()
            for _ in range(10):
                run()
            res["chunk=2^%d slots=%d %s" % (lg, ns, mode)] = round((time.perf_counter() - t0) / 10 * 1e3, 3)
print(json.dumps(res, indent=0))
