#!/usr/bin/env python
"""Times every BASELINE.json configuration's kernels on one B200 (not the headline bench --
that is bench.py; this is the per-kernel roofline table of DESIGN.md §4).

    python scripts/bench_configs.py [--steps K] [--out gpurun_out/<tag>/configs.json]

For each config: K back-to-back launches in one CUDA graph over rotating buffer sets that
together exceed the 126 MB L2, CUDA events on the launching stream, algorithmic bytes per
env-step from SURVEY.md §8d (8*W + 16 for step; reset = bytes written).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import gym_pomdp_b200 as gp  # noqa: E402

L2 = 126 * 2 ** 20


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


QUICK = False


def time_graph(fn, K, dev):
    """fn(i) enqueues launch i; returns ms per launch over K launches replayed from one graph."""
    if QUICK:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(0)
        e0.record()
        for i in range(2):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 2
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(K):
                fn(i)
        g.replay()                                   # upload + warm
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.replay()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            best = ms if best is None else min(best, ms)
    return best


def synth(env, name, B, gen, dev):
    if name.startswith("rock"):
        n, k = env.grid.x_size, env.num_rocks
        st = env.pack(torch.randint(0, n, (B,), generator=gen, device=dev), torch.randint(0, n, (B,), generator=gen, device=dev),
                      torch.randint(-1, 2, (B, k), generator=gen, device=dev))
        return st, torch.randint(0, 5 + k, (B,), generator=gen, device=dev, dtype=torch.int32)
    if name == "tag":
        st = env.pack(torch.randint(0, 29, (B,), generator=gen, device=dev), torch.randint(0, 29, (B, 1), generator=gen, device=dev))
        return st, torch.randint(0, 5, (B,), generator=gen, device=dev, dtype=torch.int32)
    if name == "tiger":
        return env.pack(torch.randint(0, 2, (B,), generator=gen, device=dev)), \
            torch.randint(0, 3, (B,), generator=gen, device=dev, dtype=torch.int32)
    if name == "network":
        return torch.randint(0, 1024, (B,), generator=gen, device=dev, dtype=torch.int32), \
            torch.randint(0, 21, (B,), generator=gen, device=dev, dtype=torch.int32)
    if name == "battleship":
        st, _ = env.init_states(B, step_ctr=int(torch.randint(1, 1 << 20, (1,)).item()))
        occ, _, _, _ = env.unpack(st)
        vis = torch.rand(occ.shape, generator=gen, device=dev) < 0.3
        rem = (occ & ~vis).reshape(B, -1).sum(1)
        vis[rem == 0] = False
        rem = (occ & ~vis).reshape(B, -1).sum(1)
        return env.pack(occ, vis, total_remaining=rem), torch.randint(0, 100, (B,), generator=gen, device=dev, dtype=torch.int32)
    raise KeyError(name)


CONFIGS = [
    # name, make-id, kwargs, log2 batch, label
    ("rock", "Rock-v0", dict(board_size=7, num_rocks=8), 20, "RockSample(7,8) B=2^20"),
    ("rock", "Rock-v0", dict(board_size=11, num_rocks=11), 22, "RockSample(11,11) B=2^22 (metric)"),
    ("rock", "Rock-v0", dict(board_size=15, num_rocks=15), 22, "RockSample(15,15) B=2^22 (one shard of 2^25/8)"),
    ("rock", "Rock-v0", dict(board_size=15, num_rocks=15), 25, "RockSample(15,15) B=2^25 (whole batch, one GPU)"),
    ("rock", "StochasticRock-v0", dict(board_size=11, num_rocks=11), 22, "StochasticRock(11,11) B=2^22"),
    ("tag", "Tag-v0", {}, 20, "Tag-v0 B=2^20"),
    ("tag", "Tag-v0", {}, 22, "Tag-v0 B=2^22"),
    ("battleship", "Battleship-v0", dict(board_size=(10, 10)), 18, "BattleShip 10x10 B=2^18"),
    ("battleship", "Battleship-v0", dict(board_size=(10, 10)), 20, "BattleShip 10x10 B=2^20"),
    ("tiger", "Tiger-v0", {}, 22, "Tiger-v0 B=2^22"),
    ("network", "Network-v0", {}, 22, "Network-v0 B=2^22"),
]


def issue_roofline(label, kernel, B, ms, sm_mhz=1965.0, sms=148):
    """For the kernels that are bound by instruction issue, not by HBM (Network's step: one Philox call per machine):
    warp instructions per launch from the committed ncu capture (profiles/issue_counts.json, keyed by config/kernel,
    counted at the batch size given there and scaled linearly) against the SMs' issue capacity -- one warp instruction
    per cycle per scheduler for the whole kernel, one per two cycles for each of the logic (ALU) and multiply-add
    (FMA) pipes."""
    try:
        counts = json.load(open(os.path.join(ROOT, "profiles", "issue_counts.json")))
    except Exception:  # noqa: BLE001
        return None
    for c in counts.get("kernels", []):
        if c["config"] in label and c["kernel"] == kernel:
            scale = B / float(c["batch"])
            cycles = ms * 1e-3 * sm_mhz * 1e6 * sms * 4              # scheduler-cycles available during the launch
            inst, alu, fma = c["inst_executed"] * scale, c["pipe_alu"] * scale, c["pipe_fma"] * scale
            return {"bound": "int-issue", "unit": "warp-inst/s", "achieved": inst / (ms * 1e-3),
                    "peak": sm_mhz * 1e6 * sms * 4, "frac": inst / cycles,
                    "alu_pipe_frac": 2 * alu / cycles, "fma_pipe_frac": 2 * fma / cycles,
                    "warp_inst_per_env_step": inst / B, "alu_inst_per_env_step": alu / B, "fma_inst_per_env_step": fma / B,
                    "sm_mhz_assumed": sm_mhz, "source": counts.get("source")}
    return None


def run_configs(dev, steps=400, only=None, kernels=("step", "step_packed", "reset", "belief_hist", "rollout"), configs=None,
                emit=None):
    """Times the named kernels of every configuration (or those whose label contains ``only``); returns the rows."""
    peak, peak_src = peak_gbs()
    rows = []

    def add(row):
        rows.append(row)
        if emit:
            emit(row)
    for name, env_id, kw, lg, label in (configs or CONFIGS):
        if only and only not in label:
            continue
        B = 1 << lg
        env = gp.make(env_id, batch_size=B, device=dev, seed=0x5EED, **kw)
        W = env.state_words
        step_bytes = 8 * W + 16
        n_sets = max(2, -(-3 * L2 // (B * step_bytes)))
        gen = torch.Generator(device=dev)
        gen.manual_seed(0x5EED)
        sets = []
        for _ in range(n_sets):
            s, a = synth(env, name, B, gen, dev)
            out = (torch.empty_like(s), torch.empty(B, dtype=torch.int32, device=dev),
                   torch.empty(B, dtype=torch.float32, device=dev), torch.empty(B, dtype=torch.int32, device=dev))
            sets.append((s, a, out))
        K = max(20, min(steps, int(steps * (1 << 22) / B)))

        if "step" in kernels:
            def step(i):
                s, a, o = sets[i % n_sets]
                env.simulate(s, a, out=o, step_ctr=i + 1)
            ms = time_graph(step, K, dev)
            gbs = B * step_bytes / (ms * 1e-3) / 1e9
            row = {"config": label, "kernel": "step", "batch": B, "state_words": W, "bytes_per_unit": step_bytes,
                   "us_per_launch": ms * 1e3, "units_per_s": B / (ms * 1e-3), "achieved_gbs": gbs, "frac_of_peak": gbs / peak,
                   "launches": K, "buffer_sets": n_sets}
            ir = issue_roofline(label, "step", B, ms)
            if ir:
                row["issue_roofline"] = ir
            if W in (1, 2) and B % 4 == 0:
                # the compute-free probe of this traffic pattern (two read + four write streams; two-word states as 256-bit
                # accesses) at this batch size: what a kernel that only moves the step's bytes takes, fixed launch cost included
                from gym_pomdp_b200 import _lib

                def probe(i):
                    s, a, o = sets[i % n_sets]
                    _lib.check(_lib.lib().pomdp_stream_probe_words(W, s.data_ptr(), a.data_ptr(), o[0].data_ptr(), o[1].data_ptr(),
                                                                   o[2].data_ptr(), o[3].data_ptr(), B,
                                                                   torch.cuda.current_stream(dev).cuda_stream), "pomdp_stream_probe_words")
                try:
                    pms = time_graph(probe, K, dev)
                    row["pattern_roof_us"] = pms * 1e3
                    row["frac_of_pattern_roof"] = pms / ms
                except Exception as e:  # noqa: BLE001
                    row["pattern_roof_error"] = repr(e)[:120]
            add(row)

        if "step_packed" in kernels and name != "battleship":
            psets = [(torch.empty_like(sets[0][0]), torch.empty(B, dtype=torch.int32, device=dev)) for _ in range(n_sets)]

            def step_packed(i):
                s, a, _ = sets[i % n_sets]
                env.simulate(s, a, out=psets[i % n_sets], step_ctr=i + 1, packed=True)
            ms = time_graph(step_packed, K, dev)
            pb = 8 * W + 8
            gbs = B * pb / (ms * 1e-3) / 1e9
            add({"config": label, "kernel": "step_packed", "batch": B, "state_words": W, "bytes_per_unit": pb,
                 "us_per_launch": ms * 1e3, "units_per_s": B / (ms * 1e-3), "achieved_gbs": gbs, "frac_of_peak": gbs / peak,
                 "note": "obs|flags|reward in one int32 stream: 8W+8 bytes per env-step"})
            del psets

        if "reset" in kernels:
            # reset: bytes written = state words + obs (+ flags for BattleShip)
            reset_bytes = 4 * W + 4 + (4 if name == "battleship" else 0)
            rsets = [(torch.empty_like(sets[0][0]), torch.empty(B, dtype=torch.int32, device=dev)) for _ in range(n_sets)]

            def reset(i):
                env.init_states(B, out=rsets[i % n_sets], step_ctr=i + 1)
            ms = time_graph(reset, max(20, K // 4), dev)
            gbs = B * reset_bytes / (ms * 1e-3) / 1e9
            add({"config": label, "kernel": "reset", "batch": B, "state_words": W, "bytes_per_unit": reset_bytes,
                 "us_per_launch": ms * 1e3, "units_per_s": B / (ms * 1e-3), "achieved_gbs": gbs, "frac_of_peak": gbs / peak})
            del rsets

        if "belief_hist" in kernels:
            hist_bytes = 4 * W

            def hist(i):
                env.belief_histogram(sets[i % n_sets][0])
            ms = time_graph(hist, max(20, K // 4), dev)
            gbs = B * hist_bytes / (ms * 1e-3) / 1e9
            add({"config": label, "kernel": "belief_hist", "batch": B, "bytes_per_unit": hist_bytes,
                 "us_per_launch": ms * 1e3, "units_per_s": B / (ms * 1e-3), "achieved_gbs": gbs, "frac_of_peak": gbs / peak})
        # fused uniform-legal rollout: states stay in registers for T steps
        if "rollout" in kernels:
            T = 32
            s0 = sets[0][0]
            ro = (torch.empty_like(s0), torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.int32, device=dev),
                  torch.empty(B, dtype=torch.int32, device=dev))
            for _ in range(2):
                env.rollout(s0, max_steps=T, out=ro, step_ctr=5)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                env.rollout(s0, max_steps=T, out=ro, step_ctr=5)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nsteps = int(ro[2].sum().item())
            add({"config": label, "kernel": "rollout(T=%d)" % T, "batch": B, "env_steps_per_launch": nsteps,
                 "us_per_launch": ms * 1e3, "units_per_s": nsteps / (ms * 1e-3),
                 "bytes_per_unit": (8 * W + 16) * B / max(nsteps, 1), "achieved_gbs": B * (8 * W + 16) / (ms * 1e-3) / 1e9,
                 "frac_of_peak": B * (8 * W + 16) / (ms * 1e-3) / 1e9 / peak,
                 "note": "compute-bound (Philox + transition logic in registers); units = env-steps actually taken"})
        del sets, env
        torch.cuda.empty_cache()
    return rows, peak, peak_src


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="substring filter on the label")
    ap.add_argument("--no-rollout", action="store_true")
    ap.add_argument("--quick", action="store_true", help="few launches per kernel (for runs under ncu)")
    ap.add_argument("--kernels", default=None, help="comma-separated subset of step,step_packed,reset,belief_hist,rollout")
    args = ap.parse_args()
    global QUICK
    QUICK = args.quick
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    kernels = ("step", "step_packed", "reset", "belief_hist") + (() if args.no_rollout else ("rollout",))
    if args.kernels:
        kernels = tuple(k for k in kernels if k in args.kernels.split(","))
    rows, peak, peak_src = run_configs(dev, args.steps, args.only, kernels, emit=lambda r: print(json.dumps(r), flush=True))
    res = {"peak_gbs": peak, "peak_source": peak_src, "gpu": torch.cuda.get_device_name(0), "rows": rows}
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
