#!/usr/bin/env python
"""bench.py -- env-steps/sec of the RockSample(11,11) step() hot path, batch 2^22 per B200.

    python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                   # CPU arm: the unmodified reference (oracle/_ref), all host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # N > 1: one rank per GPU

A "step" is ONE pass of the hot path over one batch: one ``pomdp_rock_step`` launch that
reads (state, action) for 2^22 env instances and writes (next_state, obs, reward, flags).
Env instances are independent, so N GPUs = N index shards of a global batch of N * 2^22
(weak scaling, no data-path collective; Philox is keyed by the GLOBAL env index).

Prints one JSON line (rank 0).  Keys beyond the base contract: ``roofline``, ``cpu_baseline``, ``e2e``, ``clocks``,
``gpu_launches``, ``timing`` (how the K launches were timed), ``configs`` (every other BASELINE.json configuration's
step and reset kernels: bytes per unit, microseconds per launch, roofline fraction) and ``collective`` (BASELINE
config 5: RockSample(15,15), global batch 2^25 sharded over the ranks, one step + belief histogram + NCCL all-reduce).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# rank 0 must print exactly ONE line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION in this image)
# away from it -- lower the level before anything reads it, and route fd 1 to stderr while the communicator is built
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"


class _StdoutToStderr(object):
    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False

METRIC = "env-steps/sec RockSample(11,11) batch=2^22 per B200"
UNIT = "env-steps/s"
BYTES_PER_STEP = 24          # read state 4 + action 4, write next_state 4 + obs 4 + reward 4 + flags 4
L2_BYTES = 126 * 2 ** 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000,
                    help="timed launches (default: ~175 ms of timed region, long enough to sample clocks inside it)")
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--board", type=int, default=11)
    ap.add_argument("--rocks", type=int, default=11)
    ap.add_argument("--batch", type=int, default=1 << 22, help="env instances per GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly through ctypes instead of a CUDA graph")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config kernel table and the collective row")
    ap.add_argument("--config-steps", type=int, default=200, help="launches per kernel of the per-config table")
    ap.add_argument("--min-region-ms", type=float, default=5.0,
                    help="a K-launch region shorter than this is replayed back to back and the MEDIAN replay is reported")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (for ncu --profile-from-start off; "
                         "measured to slow graph replay by ~8%, so never on for a reported number)")
    ap.add_argument("--clock-period", type=float, default=0.005, help="NVML polling period (s) during the timed region")
    return ap.parse_args()


# ------------------------------------------------------------------------- clocks ---
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            if period <= 0:
                raise RuntimeError("disabled")
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except (ValueError, IndexError):
            return local
    return local


# --------------------------------------------------------------------- reference arm ---
def run_reference(args):
    """The reference's own RockSample step() loop on the host cores: the UNMODIFIED reference package (copied by
    ``make -C oracle _ref`` into the git-ignored oracle/_ref/, which travels with the snapshot) driven through
    ``RockEnv._set_state(s); RockEnv.step(a)`` (rock.py:243-245, 123-194), every core over its own bounded sample per
    step (oracle/cpu_baseline.py).  Falls back to the oracle's Python port of the same loop when the package is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as C
    procs = os.cpu_count() or 1
    use_ref = C.reference_available()
    per_core = 1.1e4 if use_ref else 7e5            # env-steps/s/core, roughly: sizes the bounded sample
    count = max(500, min(60000 if use_ref else 400000, int(60.0 * per_core / max(1, args.steps + args.warmup))))
    arm = (C.RockReferenceArm if use_ref else C.RockCpuArm)(args.board, args.rocks, count, procs)
    for _ in range(args.warmup):
        arm.step()
    total, t = 0, 0.0
    for _ in range(args.steps):
        n, dt = arm.step()
        total += n
        t += dt
    arm.close()
    v = total / t
    what = ("the unmodified reference (oracle/_ref): RockEnv._set_state(s); RockEnv.step(a)" if use_ref
            else "oracle.pomdp_oracle.rock_step (Python port; oracle/_ref absent)")
    sample = "%d steps x %d procs x %d RockSample(%d,%d) (state, action) pairs per step, %s" % (
        args.steps, procs, count, args.board, args.rocks, what)
    cpu = {"value": v, "unit": UNIT, "cores": procs, "kind": "reference" if use_ref else "port", "sample": sample}
    if use_ref:
        one = C.time_rock_reference(args.board, args.rocks, 20000, 1)
        cpu["single_core"] = {k_: one[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
        port = C.time_rock(args.board, args.rocks, 200000, procs)
        cpu["python_port"] = {k_: port[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
    cpu["c_port"] = C.time_rock_c(args.board, args.rocks)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline_leg(args, n, k):
    """cpu_baseline of the B200 line (rank 0, N = 1): the unmodified reference on every core and on one core, the
    Python port and the C restatement beside it; about 25 s of CPU work in total."""
    from oracle import cpu_baseline as C
    procs = os.cpu_count() or 1
    keys = ("value", "unit", "cores", "kind", "sample")
    if C.reference_available():
        per_proc = max(5000, int(args.cpu_seconds * 1.0e4))
        cpu = C.time_rock_reference(n, k, per_proc, procs)
        cpu = {k_: cpu[k_] for k_ in keys}
        one = C.time_rock_reference(n, k, max(5000, per_proc // 3), 1)
        cpu["single_core"] = {k_: one[k_] for k_ in keys}
        port = C.time_rock(n, k, 400000, procs)
        cpu["python_port"] = {k_: port[k_] for k_ in keys}
    else:
        cpu = C.time_rock(n, k, max(20000, int(args.cpu_seconds * 7e5)), procs)
        cpu = {k_: cpu[k_] for k_ in keys}
    cpu["c_port"] = C.time_rock_c(n, k)          # the same algorithm as compiled C on every core (extra, not the arm)
    return cpu


def workload_config(args, n_gpus):
    return {"workload": "RockSample(%d,%d) step(): synthetic (state, action) -> (next_state, obs, reward, flags)"
                        % (args.board, args.rocks),
            "batch_per_gpu": args.batch, "global_batch": args.batch * n_gpus,
            "parallelism": "index-shard x%d (no collective)" % n_gpus,
            "state_words": 1 if args.rocks <= 11 else 2,
            "inputs": "x,y~U{0..n-1}, status~U{-1,0,1}, action~U{0..4+k}, seed 0x5EED"}


# ------------------------------------------------------------------------- B200 arm ---
def run_b200(args):
    import torch
    import torch.distributed as dist

    import gym_pomdp_b200 as gp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- gym_pomdp_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()                      # builds the communicator (and prints NCCL's banner, if any) now
    n_gpus = world
    B, n, k = args.batch, args.board, args.rocks
    K, W = args.steps, args.warmup

    env = gp.make("Rock-v0", board_size=n, num_rocks=k, batch_size=B, device=dev, seed=0x5EED, global_offset=rank * B)
    words = env.state_words
    bytes_per_step = 8 * words + 16
    set_bytes = B * bytes_per_step
    # rotate through enough independent buffer sets that a launch never finds its inputs in L2
    n_sets = max(2, -(-4 * L2_BYTES // set_bytes))
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x5EED + rank)
    sets = []
    for _ in range(n_sets):
        x = torch.randint(0, n, (B,), generator=gen, device=dev)
        y = torch.randint(0, n, (B,), generator=gen, device=dev)
        status = torch.randint(-1, 2, (B, k), generator=gen, device=dev)
        action = torch.randint(0, 5 + k, (B,), generator=gen, device=dev, dtype=torch.int32)
        state = env.pack(x, y, status)
        del x, y, status
        out = (torch.empty_like(state), torch.empty(B, dtype=torch.int32, device=dev),
               torch.empty(B, dtype=torch.float32, device=dev), torch.empty(B, dtype=torch.int32, device=dev))
        sets.append((state, action, out))

    def launch(i):
        s, a, o = sets[i % n_sets]
        env.simulate(s, a, out=o, step_ctr=i + 1)

    # warm-up (also loads the module, sizes the grid)
    for i in range(max(W, 3) if args.no_graph else 3):
        launch(i)
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(dev)
    graph_w = graph_t = None
    if not args.no_graph:
        # One CUDA graph holding exactly K step launches (and one with W for the warm-up):
        # removes the Python/ctypes launch path from a ~20 us kernel's critical path.
        with torch.cuda.stream(stream):
            graph_w = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_w, stream=stream):
                for i in range(max(W, 3)):
                    launch(i)
            graph_t = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_t, stream=stream):
                for i in range(K):
                    launch(i)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            if graph_t is not None:
                graph_t.replay()
            else:
                for i in range(K):
                    launch(i)
            e1.record(stream)
        return e0, e1

    with torch.cuda.stream(stream):
        if graph_w is not None:
            graph_w.replay()
            # the first launch of a graph exec uploads its 2000 nodes to the device (~1.5 us per node,
            # measured): do that outside the timed region, as cudaGraphUpload would
            graph_t.replay()
    barrier()
    sampler = ClockSampler(physical_gpu_index(local), period=args.clock_period)
    if args.clock_period > 0:
        sampler.start()
    if args.profiler_range:
        torch.cuda.profiler.start()     # ncu --profile-from-start off: list only the timed region's launches
    e0, e1 = timed_region()
    torch.cuda.synchronize()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    barrier()
    ms = e0.elapsed_time(e1)
    timing = {"mode": "one region of K launches", "region_ms": ms, "replays": 1}
    clocks_note = "sampled during the timed region"
    if graph_t is not None and ms < args.min_region_ms and not args.profiler_range:
        # K launches of a ~17 us kernel are a few hundred microseconds: too short for a stable number (and for NVML to
        # see).  Replay the same K-launch graph back to back, every replay bracketed by its own pair of events, and
        # report the MEDIAN replay (steps stays K; every replay is the same K launches over the same rotating buffers).
        R = int(min(2000, max(9, 150.0 / max(ms, 1e-3))))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(R)]
        barrier()
        with torch.cuda.stream(stream):
            for a_, b_ in evs:
                a_.record(stream)
                graph_t.replay()
                b_.record(stream)
        torch.cuda.synchronize()
        barrier()
        reps = sorted(a_.elapsed_time(b_) for a_, b_ in evs)
        timing = {"mode": "median of R back-to-back replays of the K-launch graph (the single region is shorter than "
                          "%.1f ms)" % args.min_region_ms, "first_region_ms": ms, "replays": R,
                  "median_ms": reps[R // 2], "min_ms": reps[0], "p90_ms": reps[(9 * R) // 10]}
        ms = reps[R // 2]
        clocks_note = "sampled over the timed region and its %d timed replays (%.0f ms)" % (R, sum(reps))
    elif args.clock_period > 0 and len(sampler.samples) < 5:
        # the region is shorter than a few NVML polls: repeat the identical work (untimed) while sampling
        t_end = time.time() + 0.6
        while time.time() < t_end:
            timed_region()
            torch.cuda.synchronize()
        clocks_note = "timed region %.1f ms is shorter than 5 NVML polls; sampled over untimed repeats of it" % ms
    if args.clock_period > 0:
        sampler.stop()
    clocks = sampler.summary()
    clocks["note"] = clocks_note

    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_gpus * B * K / (ms * 1e-3)

    # ---- e2e: host buffers in, host buffers out, through the public API ---------------
    # env.simulate_host(packed=True) = ONE C-ABI call, pomdp_step_packed_host (include/pomdp_b200.h): pinned host
    # (state, action) -> the library's chunked H2D / kernel / D2H pipeline on three streams -> pinned host
    # (next_state, result) with result = obs | flags << 8 | reward << 16.  Timed beside it: the same pipeline driven
    # from Python through torch (e2e.python_pipeline), the four-array result (e2e.unpacked, 16 B/env back) and the
    # zero-copy launch (e2e.zero_copy).
    E = max(1, args.e2e_steps)
    pin = dict(device="cpu", pin_memory=True)
    s0, a0, _ = sets[0]
    h_state, h_action = s0.cpu().pin_memory(), a0.cpu().pin_memory()
    h_out = (torch.empty(s0.shape, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.int32, **pin),
             torch.empty(B, dtype=torch.float32, **pin), torch.empty(B, dtype=torch.int32, **pin))
    h_packed = (torch.empty(s0.shape, dtype=torch.int32, **pin), torch.empty(B, dtype=torch.int32, **pin))

    def time_e2e(packed, pipeline=None):
        out = h_packed if packed else h_out
        for i in range(5):
            env.simulate_host(h_state, h_action, out, step_ctr=i + 1, packed=packed, pipeline=pipeline)
        barrier()
        t0 = time.perf_counter()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for i in range(E):
            env.simulate_host(h_state, h_action, out, step_ctr=i + 1, packed=packed, pipeline=pipeline)
        ee1.record()
        torch.cuda.synchronize()
        ms_ = ee0.elapsed_time(ee1)
        wall = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([ms_], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_, wall

    e2e_un_ms, _ = time_e2e(False)
    e2e_py_ms, _ = time_e2e(True, "python")
    h_packed[0].zero_(); h_packed[1].zero_()
    e2e_ms, e2e_wall = time_e2e(True)                       # the C-ABI host call; its results are checked below
    # zero-copy variant: one launch, the kernel itself reads / writes the pinned host buffers over PCIe
    zc_ms = None
    try:
        for i in range(3):
            env.simulate_host(h_state, h_action, h_packed, step_ctr=i + 1, packed=True, zero_copy=True)
        barrier()
        z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        z0.record()
        for i in range(E):
            env.simulate_host(h_state, h_action, h_packed, step_ctr=i + 1, packed=True, zero_copy=True)
        z1.record()
        torch.cuda.synchronize()
        zc_ms = z0.elapsed_time(z1)
        if world > 1:
            t = torch.tensor([zc_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            zc_ms = float(t.item())
    except Exception:  # noqa: BLE001 - platforms without mapped pinned memory
        zc_ms = None
    e2e_value = n_gpus * B * E / (e2e_ms * 1e-3)
    # what the link gives: the same bytes each way (pinned host <-> device), both directions at once on two streams, no
    # kernel and no dependency between them -- the ceiling of any host-buffer call on this box with this many ranks
    cs_in, cs_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_in = (torch.empty_like(s0), torch.empty_like(a0))
    d_out = (torch.empty_like(s0), torch.empty(B, dtype=torch.int32, device=dev))

    def copies():
        with torch.cuda.stream(cs_in):
            d_in[0].copy_(h_state, non_blocking=True)
            d_in[1].copy_(h_action, non_blocking=True)
        with torch.cuda.stream(cs_out):
            h_packed[0].copy_(d_out[0], non_blocking=True)
            h_packed[1].copy_(d_out[1], non_blocking=True)
    saved = (h_packed[0].clone(), h_packed[1].clone())
    for _ in range(3):
        copies()
    torch.cuda.synchronize()
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    cs_in.wait_event(c0); cs_out.wait_event(c0)
    for _ in range(E):
        copies()
    torch.cuda.current_stream().wait_stream(cs_in); torch.cuda.current_stream().wait_stream(cs_out)
    c1.record()
    torch.cuda.synchronize()
    ceil_ms = c0.elapsed_time(c1)
    if world > 1:
        t = torch.tensor([ceil_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ceil_ms = float(t.item())
    h_packed[0].copy_(saved[0]); h_packed[1].copy_(saved[1])
    del d_in, d_out, saved
    # sanity: the host results of the last e2e steps equal the device path on the same inputs
    chk = env.simulate(s0, a0, step_ctr=E)
    assert torch.equal(chk[1].cpu(), h_out[1]) and torch.equal(chk[0].cpu(), h_out[0]), "e2e result != device result"
    p_ob, p_rw, p_fl = env.unpack_result(h_packed[1])
    assert torch.equal(h_packed[0], h_out[0]) and torch.equal(p_ob, h_out[1]) and torch.equal(p_rw, h_out[2]) \
        and torch.equal(p_fl, h_out[3]), "packed e2e result != unpacked e2e result"

    collective = None
    if not args.no_configs:
        collective = collective_row(gp, torch, dist, dev, rank, world)
    if rank != 0:
        if world > 1:
            dist.barrier()               # rank 0 is still timing the per-config table
            dist.destroy_process_group()
        return

    peaks, peak_src = None, "fallback 6650 GB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:  # noqa: BLE001
        peak = 6650.0
    per_launch_ms = ms / K
    achieved = B * bytes_per_step / (per_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "traffic_l2": ncu_traffic("l2_bytes_from_sms_per_launch"),
                "traffic_source": "NOT measured in this run: dram__bytes_read+write and lts__t_sectors_srcunit_tex of one "
                                  "profiled launch, committed ncu capture %s (profiles/rock_step_ncu_summary.json)"
                                  % (ncu_traffic("tag") or "?"),
                "kernel": "pomdp_step_kernel<RockEnv%d,true>" % words,
                "algorithmic_bytes_per_launch": B * bytes_per_step, "avg_launch_us": per_launch_ms * 1e3,
                "peak_source": peak_src,
                "nominal_peak": 8000.0, "frac_of_nominal": achieved / 8000.0,     # both peaks stated
                "pattern_roof": live_pattern_roof(torch, env, sets, n_sets, B, K, stream, peak, per_launch_ms)
                if (words == 1 and B % 4 == 0) else (pattern_roof(22) if B == 1 << 22 else None),
                "read_only_frac": (B * (4 * words + 4) / (per_launch_ms * 1e-3) / 1e9) / peak}

    cpu = None
    if not args.no_cpu:
        cpu = cpu_baseline_leg(args, n, k)
    configs = None
    if not args.no_configs:
        configs = config_rows(dev, args.config_steps)

    cfg = workload_config(args, n_gpus)
    cfg["l2"] = "%d rotating buffer sets of %.0f MB (%.0f MB total) > 126 MB L2; no flush needed" % (
        n_sets, set_bytes / 1e6, n_sets * set_bytes / 1e6)
    cfg["launch"] = "eager ctypes launches" if args.no_graph else \
        "one CUDA graph of K pomdp_rock_step launches (replayed once untimed first: graph upload)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": cfg,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (4 * words + 4),
                "d2h_bytes_per_step": B * (4 * words + 4), "steps": E, "ms_per_step": e2e_ms / E,
                "wall_ms_per_step": e2e_wall / E,
                "ceiling_ms": ceil_ms / E, "frac_of_ceiling": (ceil_ms / E) / (e2e_ms / E),
                "ceiling": "pure pinned H2D + D2H copies of the same bytes (%.1f MB each way per rank), both directions at once, "
                           "no kernel, %d rank(s) at once: what the host links give this call" % (B * (4 * words + 4) / 1e6, world),
                "path": "env.simulate_host(packed=True) = one C-ABI call pomdp_step_packed_host: pinned host (state, action) -> "
                        "chunks of 2^20 envs, H2D / pomdp_rock_step_packed / D2H on the library's three streams -> pinned host "
                        "(next_state, result = obs | flags << 8 | reward << 16)",
                "python_pipeline": {"value": n_gpus * B * E / (e2e_py_ms * 1e-3), "ms_per_step": e2e_py_ms / E,
                                    "path": "the same pipeline issued from Python through torch, chunks of 2^20 envs"},
                "unpacked": {"value": n_gpus * B * E / (e2e_un_ms * 1e-3), "ms_per_step": e2e_un_ms / E,
                             "d2h_bytes_per_step": B * (4 * words + 12),
                             "path": "env.simulate_host: four result arrays (next_state, obs, reward, flags) back"},
                "zero_copy": None if zc_ms is None else {
                    "value": n_gpus * B * E / (zc_ms * 1e-3), "ms_per_step": zc_ms / E,
                    "path": "env.simulate_host(packed=True, zero_copy=True): one launch, the kernel loads from and stores to "
                            "the pinned host buffers directly (no staging copies)"}},
        "clocks": clocks, "gpu_launches": K,
        "timing": timing, "configs": configs, "collective": collective,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config_rows(dev, steps):
    """Every other BASELINE.json configuration's step and reset kernels (and Network's, the one integer-issue-bound step),
    timed like the headline: K launches in one CUDA graph over rotating buffer sets larger than L2 (scripts/bench_configs.py
    holds the code; `python scripts/bench_configs.py` prints the long table with every kernel)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(ROOT, "scripts", "bench_configs.py"))
    bc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bc)
    want = ["RockSample(7,8) B=2^20", "Tag-v0 B=2^20", "BattleShip 10x10 B=2^18", "RockSample(15,15) B=2^22", "Network-v0 B=2^22",
            "RockSample(11,11) B=2^22"]
    cfgs = [c for c in bc.CONFIGS if any(c[4].startswith(w) for w in want)]
    rows, peak, _ = bc.run_configs(dev, steps, None, ("step", "reset"), configs=cfgs)
    out = []
    for r in rows:
        if r["config"].startswith("RockSample(11,11)") and r["kernel"] == "step":
            continue                                # the headline itself
        o = {"workload": r["config"], "kernel": r["kernel"], "bytes_per_unit": r["bytes_per_unit"],
             "us_per_launch": r["us_per_launch"], "units_per_s": r["units_per_s"], "roofline_frac": r["frac_of_peak"],
             "bound": "hbm"}
        if "issue_roofline" in r:
            o["issue_roofline"] = r["issue_roofline"]
        if "pattern_roof_us" in r:          # the compute-free probe of the step's six streams at this batch size, same run
            o["pattern_roof_us"], o["frac_of_pattern_roof"] = r["pattern_roof_us"], r["frac_of_pattern_roof"]
        out.append(o)
    return out


def collective_row(gp, torch, dist, dev, rank, world):
    """BASELINE.json config 5: RockSample(15,15), global batch 2^25 sharded over the ranks by index, one step, then the
    belief histogram (271 int64 bins) summed over the ranks by NCCL -- the only collective on the path.  Every rank
    takes part; the reduced histogram is checked against the gathered per-rank ones."""
    G = 1 << 25
    B = G // world
    env = gp.make("Rock-v0", board_size=15, num_rocks=15, batch_size=B, device=dev, seed=0x5EED, global_offset=rank * B)
    state, _ = env.init_states(B, step_ctr=1)
    action = env.sample_legal_actions(state, step_ctr=2)
    nxt = env.simulate(state, action, step_ctr=2)[0]
    del state, action
    local = env.belief_histogram(nxt)
    red = env.belief_histogram(nxt, all_reduce=True)
    ok = True
    if world > 1:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        ok = bool(torch.equal(red, torch.stack(parts).sum(0)))
    ok = ok and int(red[15:].sum().item()) == G                 # every particle sits in exactly one agent-cell bin
    for _ in range(3):
        env.belief_histogram(nxt, all_reduce=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        env.belief_histogram(nxt, all_reduce=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(reps):
        env.belief_histogram(nxt)
    h1.record()
    torch.cuda.synchronize()
    hist_us = h0.elapsed_time(h1) * 1e3 / reps
    # the same reduction made inside the histogram kernel over NVLink peer memory (pomdp_belief_hist_allreduce)
    fused_us, fused_ok, fused_err = None, None, None
    if world > 1:
        try:
            fused = env.belief_histogram(nxt, all_reduce="fused")
            fused_ok = bool(torch.equal(fused, red))
            for _ in range(3):
                fused_ok = fused_ok and bool(torch.equal(env.belief_histogram(nxt, all_reduce="fused"), red))
            torch.cuda.synchronize()
            dist.barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(reps):
                env.belief_histogram(nxt, all_reduce="fused")
            f1.record()
            torch.cuda.synchronize()
            fused_us = f0.elapsed_time(f1) * 1e3 / reps
        except Exception as e:  # noqa: BLE001
            fused_err = repr(e)[:200]
    # the same three variants with the interpreter out of the way: 20 calls captured in one CUDA graph each (eager numbers
    # of ~10 us kernels are set by launch overhead and host jitter)
    graph_us = None
    if world > 1 and fused_err is None:
        try:
            graph_us = {}
            for name, fn in (("hist_only", lambda: env.belief_histogram(nxt)),
                             ("hist_plus_allreduce", lambda: env.belief_histogram(nxt, all_reduce=True)),
                             ("fused", lambda: env.belief_histogram(nxt, all_reduce="fused"))):
                gs = torch.cuda.Stream(dev)
                with torch.cuda.stream(gs):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=gs):
                        for _ in range(reps):
                            fn()
                    g.replay()
                    torch.cuda.synchronize()
                    dist.barrier()
                    ts = []
                    for _ in range(5):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(gs)
                        g.replay()
                        b.record(gs)
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b) * 1e3 / reps)
                tt = torch.tensor([sorted(ts)[2]], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                graph_us[name] = float(tt.item())
                del g
        except Exception as e:  # noqa: BLE001
            graph_us = {"error": repr(e)[:200]}
    if world > 1:
        t = torch.tensor([us, hist_us, fused_us if fused_us is not None else -1.0, 1.0 if fused_ok else 0.0], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us, hist_us, fmax, _ = (float(v) for v in t.tolist())
        tmin = torch.tensor([1.0 if fused_ok else 0.0, 0.0 if fused_us is None else 1.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        fused_ok = bool(tmin[0].item() == 1.0)
        fused_us = fmax if tmin[1].item() == 1.0 else None
    # Config 5's whole step -- transition of the shard, counts of the next states, sum over the ranks -- as separate
    # launches and as ONE kernel (pomdp_rock_step_hist: the step kernel counts the next states in its epilogue and the
    # all-reduce rides in the same launch over NVLink peer memory).  Graph-captured, median of 5 replays, max over ranks.
    pipeline = None
    try:
        state2, _ = env.init_states(B, step_ctr=1)
        action2 = env.sample_legal_actions(state2, step_ctr=2)
        sout = (nxt, torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.float32, device=dev),
                torch.empty(B, dtype=torch.int32, device=dev))
        fused_mode = "fused" if (world > 1 and fused_err is None) else False

        def separate(all_reduce):
            env.simulate(state2, action2, out=sout, step_ctr=2)
            return env.belief_histogram(sout[0], all_reduce=all_reduce)

        def graph_us_of(fn, n_calls=10):
            gs = torch.cuda.Stream(dev)
            with torch.cuda.stream(gs):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=gs):
                    for _ in range(n_calls):
                        fn()
                g.replay()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                ts = []
                for _ in range(5):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(gs)
                    g.replay()
                    b.record(gs)
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e3 / n_calls)
            tt = torch.tensor([sorted(ts)[2]], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            del g
            return float(tt.item())

        want = separate(world > 1)
        got = env.simulate_hist(state2, action2, out=sout, step_ctr=2, all_reduce=fused_mode)[4]
        pipe_ok = bool(torch.equal(got, want)) and bool(torch.equal(red, want))
        pipeline = {"step_only_us": graph_us_of(lambda: env.simulate(state2, action2, out=sout, step_ctr=2))}
        if world > 1:
            pipeline["step_hist_allreduce_three_launches_us"] = graph_us_of(lambda: separate(True))
            if fused_mode:
                pipeline["step_fusedhist_two_launches_us"] = graph_us_of(lambda: separate("fused"))
        else:
            pipeline["step_hist_two_launches_us"] = graph_us_of(lambda: separate(False))
        pipeline["step_hist%s_one_launch_us" % ("_allreduce" if fused_mode else "")] = graph_us_of(
            lambda: env.simulate_hist(state2, action2, out=sout, step_ctr=2, all_reduce=fused_mode))
        if world > 1:
            tk = torch.tensor([1.0 if pipe_ok else 0.0], device=dev, dtype=torch.float64)
            dist.all_reduce(tk, op=dist.ReduceOp.MIN)
            pipe_ok = bool(tk.item() == 1.0)
        pipeline["equal_counts"] = pipe_ok
        pipeline["what"] = ("simulate_hist = pomdp_rock_step_hist: the step kernel counts next_state in its epilogue (it is never "
                            "read back from HBM) and hands the counts out -- with all_reduce='fused' summed over the ranks in the "
                            "same launch; 10 calls per CUDA graph, median of 5 replays, max over ranks")
        del state2, action2, sout
    except Exception as e:  # noqa: BLE001
        pipeline = {"error": repr(e)[:200]}
    del nxt, env
    torch.cuda.empty_cache()
    row = {"workload": "RockSample(15,15) global batch 2^25 over %d rank(s): belief histogram%s" % (
               world, " + ncclAllReduce(sum)" if world > 1 else " (one rank: no collective)"),
           "batch_per_gpu": B, "bins": int(red.numel()), "hist_plus_allreduce_us": us, "hist_only_us": hist_us, "ok": ok,
           "note": "eager launches (self-cleaning histogram kernel + all-reduce), CUDA events, max over ranks"}
    if world > 1:
        row["fused"] = {"hist_allreduce_us": fused_us, "equals_nccl": fused_ok, "error": fused_err,
                        "what": "pomdp_belief_hist_allreduce: ONE kernel -- the last CTA of every rank adds the rank's counts into "
                                "every rank's buffer over NVLink peer memory, signals its arrival to every peer, waits for "
                                "theirs and writes the global counts out; no zero-fill, no NCCL call, no separate barrier"}
        row["graph_us"] = graph_us
    row["step_pipeline"] = pipeline
    return row


def ncu_traffic(key="dram_bytes_per_launch"):
    """Per-launch traffic of the step kernel from the committed ncu capture (profiles/rock_step_ncu_summary.json), if any:
    dram__bytes_read + dram__bytes_write (a single profiled launch leaves most of its 67 MB of writes in the 126 MB L2,
    so this undercounts the writes), or the bytes the SMs moved through L2 (= the algorithmic bytes when nothing is re-read)."""
    p = os.path.join(ROOT, "profiles", "rock_step_ncu_summary.json")
    try:
        with open(p) as f:
            return json.load(f).get(key)
    except Exception:  # noqa: BLE001
        return None


def live_pattern_roof(torch, env, sets, n_sets, B, K, stream, peak, step_ms):
    """Measured in THIS run: pomdp_stream_probe -- the step kernel's six memory streams with no table, no draws and no
    transition -- timed exactly like the step (the same rotating buffer sets, one CUDA graph, PDL, CUDA events on the
    launching stream).  `step_vs_roof` = probe time / step time: 1.0 means the step adds nothing to its own memory traffic."""
    import statistics as st
    from gym_pomdp_b200 import _lib
    L = _lib.lib()
    Kp = max(20, min(K, 2000))

    def launch(i):
        s, a, o = sets[i % n_sets]
        _lib.check(L.pomdp_stream_probe(s.data_ptr(), a.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                        o[3].data_ptr(), B, torch.cuda.current_stream(s.device).cuda_stream), "pomdp_stream_probe")
    try:
        for i in range(3):
            launch(i)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for i in range(Kp):
                    launch(i)
            g.replay()                                   # graph upload
            torch.cuda.synchronize()
            times = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                g.replay()
                e1.record(stream)
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) / Kp)
        ms = st.median(times)
        return {"us_per_launch": ms * 1e3, "gbs": B * 24 / (ms * 1e-3) / 1e9, "frac_of_peak": B * 24 / (ms * 1e-3) / 1e9 / peak,
                "step_vs_roof": ms / step_ms, "launches": Kp, "replays": 5,
                "source": "measured in this run: pomdp_stream_probe (compute-free kernel with the step's two read + four write "
                          "streams, same grid, graph + PDL), median of 5 graph replays"}
    except Exception as e:  # noqa: BLE001
        r = pattern_roof(22) if B == 1 << 22 else None
        if r:
            r["live_error"] = repr(e)
        return r


def pattern_roof(log2_batch=22):
    """The compute-free probe of the step's traffic pattern (scripts/exp_stream_probe.cu: two read + four write streams, 24 B
    per env, the product's grid, PDL + graph) from the committed B200 run: what this read:write mix reaches on the part."""
    import re
    name = "r04k_stream_probe.log"
    try:
        for line in open(os.path.join(ROOT, "profiles", name)):
            m = re.match(r"n=2\^(\d+) .*graph\+PDL ([0-9.]+) us", line)
            if m and int(m.group(1)) == log2_batch:
                return {"us_per_launch": float(m.group(2)), "source": "NOT measured in this run: profiles/" + name +
                        " (compute-free kernel with the same six streams, graph + PDL)"}
    except Exception:  # noqa: BLE001
        pass
    return None


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
