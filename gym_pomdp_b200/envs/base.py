"""Common host-side machinery of the batched POMDP environments.

The reference's envs are single-instance Python objects with the old-gym protocol
(``reset() -> ob``; ``step(a) -> (ob, reward, done, {"state": s})``) plus the underscore
hooks planners use (``_set_state``, ``_get_init_state``, ``_generate_legal``,
``_compute_prob`` ...).  Here one object holds ``batch_size`` independent
instances as packed int32 words in a torch tensor on a B200 and every call is one CUDA
kernel reached through the ctypes C ABI (gym_pomdp_b200/_lib.py).

Two calling modes:

* ``batch_size=None`` (default, what ``gym.make(id)`` callers get): one instance, Python
  scalars in and out, ``info["state"]`` / ``_set_state`` in the reference's own formats,
  and the reference's ``assert``/``IndexError`` behaviour on bad calls -- a drop-in.
* ``batch_size=B``: tensors in and out (``obs`` int32[B], ``reward`` float32[B], ``done``
  bool[B]); ``info["state"]`` is the packed state tensor int32[B, words] and
  ``info["flags"]`` carries the per-instance error bits that replace the asserts.

``simulate(state, action)`` is the functional generative model G(s, a) -> (s', o, r, flags)
on caller-owned particle tensors: what a POMCP / particle-filter caller does with
``_set_state(s); step(a)`` in a Python loop, as one launch.
"""
import ctypes
import os

import torch

from .. import _lib


def _env_base():
    """The reference's envs subclass gym.Env (rock.py:5, 96) and speak the OLD gym protocol (reset() -> ob; four values from
    step).  Subclass gym.Env only where that protocol is gym's own (gym < 0.26); newer gym / gymnasium get an adapter
    around these classes instead (registration.NewApiAdapter), because their wrappers call reset(seed=, options=)."""
    try:  # pragma: no cover - gym is not in the build image
        import gym
        from ..registration import uses_new_api
        if not uses_new_api("gym", gym):
            return gym.Env
    except Exception:  # noqa: BLE001
        pass
    return object


_EnvBase = _env_base()


def _fresh_seed():
    return int.from_bytes(os.urandom(8), "little")


def _as_device(device):
    return torch.device(device) if not isinstance(device, torch.device) else device


class BatchedPomdpEnv(_EnvBase):
    metadata = {"render.modes": ["ansi"]}
    kind = -1            # POMDP_KIND_* for the belief histogram
    state_words = 1      # int32 words per instance

    def __init__(self, batch_size=None, device="cuda", seed=None, global_offset=0):
        self._scalar = batch_size is None
        self.batch_size = 1 if self._scalar else int(batch_size)
        if self.batch_size < 0:
            raise ValueError("batch_size must be >= 0")
        self.device = _as_device(device)
        _lib.lib()  # fail loudly if the CUDA library has not been built
        if self.device.type != "cuda" and not _lib.is_hostsim():
            raise RuntimeError("gym_pomdp_b200 runs on CUDA devices only (no CPU path); got device=%r" % (device,))
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # The reference draws from numpy's unseeded global RNG, so two env objects are independent; a fixed default seed
        # would make every default-constructed env replay the same rock layouts and sensor noise.  seed=None (the default)
        # therefore draws a fresh 64-bit key; pass a seed (or call seed()) for reproducible runs.
        self._seed = _fresh_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        self._step_ctr = 0
        self.global_offset = int(global_offset)   # index of instance 0 in the global batch (multi-GPU shards)
        self.state = None
        self.flags = None
        self.done = False if self._scalar else None
        self.last_action = None

    # ------------------------------------------------------------------ plumbing ----
    def _stream(self):
        return _lib.stream_handle(self.device)

    def _next_ctr(self):
        self._step_ctr = (self._step_ctr + 1) & 0xFFFFFFFF
        return self._step_ctr

    def _empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _guard(self):
        return torch.cuda.device(self.device) if self.device.type == "cuda" else _NullCtx()

    # subclasses: raw C calls on tensors --------------------------------------------
    def _c_step(self, state, action, next_state, obs, reward, flags, n, ctr):
        raise NotImplementedError

    def _c_reset(self, state, obs, mask, n, ctr):
        raise NotImplementedError

    _abi = None          # "rock", "tag", ...: prefix of the env's C entry points

    def _c_head(self):
        """Leading arguments of the env's C entry points: (params,) or (params, d_table)."""
        return (ctypes.byref(self._params),)

    def _c_policy(self, state, action, n, ctr):
        fn = getattr(_lib.lib(), "pomdp_%s_policy" % self._abi)
        _lib.check(fn(*self._c_head(), _lib.ptr(state), _lib.ptr(action), n, self.global_offset, self._seed, ctr,
                      self._stream()), "pomdp_%s_policy" % self._abi)

    def _c_rollout(self, state, final_state, ret, steps, flags, n, ctr, max_steps, discount, first_action=None):
        fn = getattr(_lib.lib(), "pomdp_%s_rollout" % self._abi)
        _lib.check(fn(*self._c_head(), _lib.ptr(state), _lib.ptr(first_action), _lib.ptr(final_state), _lib.ptr(ret), _lib.ptr(steps),
                      _lib.ptr(flags), n, self.global_offset, self._seed, ctr, int(max_steps), float(discount),
                      self._stream()), "pomdp_%s_rollout" % self._abi)

    def _c_obs_prob(self, next_state, action, obs, prob, n, *extra):
        fn = getattr(_lib.lib(), "pomdp_%s_obs_prob" % self._abi)
        _lib.check(fn(*self._c_query_head(), _lib.ptr(next_state), _lib.ptr(action), _lib.ptr(obs), _lib.ptr(prob), n, *extra,
                      self._stream()), "pomdp_%s_obs_prob" % self._abi)

    def _c_query_head(self):
        """Leading arguments of the table-free query entry points (obs_prob / legal_mask)."""
        return (ctypes.byref(self._params),)

    def observation_prob(self, action, next_state, ob, *extra):
        """Batched ``_compute_prob(action, next_state, ob)``: float64[n], one kernel (particle reweighting;
        SURVEY.md §8f rank 2).  ``next_state`` is the packed post-step state."""
        n = next_state.shape[0]
        dev = next_state.device
        action = torch.as_tensor(action, device=dev).to(torch.int32).expand(n).contiguous()
        ob = torch.as_tensor(ob, device=dev).to(torch.int32).expand(n).contiguous()
        prob = torch.empty(n, dtype=torch.float64, device=dev)
        with self._guard():
            self._c_obs_prob(next_state.contiguous(), action, ob, prob, n, *extra)
        return prob

    def legal_mask_words(self, state=None):
        """``_generate_legal()`` for every particle as bit masks over action ids: int32[n, ceil(n_actions / 32)]
        (SURVEY.md §8f rank 3), one kernel."""
        state = self.state if state is None else state
        n = state.shape[0]
        words = (self.action_space.n + 31) // 32
        mask = torch.empty((n, words), dtype=torch.int32, device=state.device)
        fn = getattr(_lib.lib(), "pomdp_%s_legal_mask" % self._abi)
        with self._guard():
            _lib.check(fn(*self._c_query_head(), _lib.ptr(state.contiguous()), _lib.ptr(mask), n, self._stream()),
                       "pomdp_%s_legal_mask" % self._abi)
        return mask

    def legal_mask(self, state=None):
        """bool[n, n_actions] view of ``legal_mask_words``"""
        w = self.legal_mask_words(state).to(torch.int64) & 0xFFFFFFFF
        a = torch.arange(self.action_space.n, device=w.device)
        return ((w[:, a // 32] >> (a % 32)) & 1).bool()

    # subclasses: scalar-mode conversions -------------------------------------------
    def _state_to_ref(self, words):
        raise NotImplementedError

    def _state_from_ref(self, ref_state):
        raise NotImplementedError

    def _reward_to_py(self, reward, action):
        return float(reward)

    def _raise_for_flags(self, flags):
        if flags & _lib.FLAG_BAD_STATE:
            raise AssertionError("state outside the environment's domain")

    # -------------------------------------------------------- functional interface ---
    _reward_unit = 1          # what one unit of a packed result's reward field is worth (Network: tenths)

    def _c_step_packed(self, state, action, next_state, result, n, ctr):
        fn = getattr(_lib.lib(), "pomdp_%s_step_packed" % self._abi, None)
        if fn is None:
            raise NotImplementedError("%s has no packed step (its state words dominate the traffic)" % type(self).__name__)
        _lib.check(fn(*self._c_head(), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state), _lib.ptr(result), n,
                      self.global_offset, self._seed, ctr, self._stream()), "pomdp_%s_step_packed" % self._abi)

    def unpack_result(self, result):
        """Packed result words (include/pomdp_b200.h) -> (obs int32, reward float32, flags int32), equal to what
        the unpacked ``simulate`` returns.  Works on CPU and CUDA tensors."""
        obs = result & 0xFF
        flags = (result >> 8) & 0xFF
        units = result >> 16                              # arithmetic shift: signed 16-bit field
        if self._reward_unit == 1:
            reward = units.to(torch.float32)
        else:
            reward = (units.to(torch.float64) / self._reward_unit).to(torch.float32)
        return obs, reward, flags

    def simulate(self, state, action, out=None, step_ctr=None, packed=False):
        """G(s, a): one transition for every particle.

        The draws of a call are a pure function of (seed, global env index, step_ctr): two calls with the same explicit
        ``step_ctr`` repeat each other's randomness.  Leave ``step_ctr`` unset (the env's own counter advances) unless the
        repetition is wanted (tests, common random numbers).

        state int32[n, words] (or [n] when words == 1), action int32[n].  Returns
        (next_state, obs, reward, flags); ``out`` may supply those four tensors
        (``out[0]`` may be ``state`` itself for an in-place step).  With ``packed=True`` the
        result is (next_state, result) with obs | flags << 8 | reward_units << 16 in one int32
        stream (``unpack_result`` decodes it): a third less device traffic, half the bytes to
        fetch for a host caller.
        """
        n = action.shape[0]
        if packed:
            if out is None:
                out = (torch.empty_like(state), self._empty((n,), torch.int32))
            ctr = self._next_ctr() if step_ctr is None else int(step_ctr)
            with self._guard():
                self._c_step_packed(state, action, out[0], out[1], n, ctr)
            return out[0], out[1]
        if out is None:
            out = (torch.empty_like(state), self._empty((n,), torch.int32), self._empty((n,), torch.float32),
                   self._empty((n,), torch.int32))
        next_state, obs, reward, flags = out
        ctr = self._next_ctr() if step_ctr is None else int(step_ctr)
        with self._guard():
            self._c_step(state, action, next_state, obs, reward, flags, n, ctr)
        return next_state, obs, reward, flags

    def simulate_hist(self, state, action, out=None, step_ctr=None, all_reduce=False, hist_out=None):
        """``simulate`` and ``belief_histogram(next_state)`` in ONE kernel (``pomdp_E_step_hist``): the next states are
        counted while they are still in registers, so the particle set is not read back from HBM by a second kernel
        ("accumulated in the step kernel epilogue").  ``all_reduce`` as in ``belief_histogram``: False
        -- this shard's counts; True -- summed over the ranks by NCCL afterwards; ``"fused"`` -- summed INSIDE the same
        launch over NVLink / NVSwitch peer memory (step + histogram + all-reduce: one kernel, replayable from a CUDA
        graph).  Returns (next_state, obs, reward, flags, counts).  BattleShip's step kernel (TMA board tiles) has no
        histogram epilogue: there the two kernels run back to back."""
        n = action.shape[0]
        if out is None:
            out = (torch.empty_like(state), self._empty((n,), torch.int32), self._empty((n,), torch.float32),
                   self._empty((n,), torch.int32))
        next_state, obs, reward, flags = out
        if not hasattr(self, "_c_step_hist"):
            self.simulate(state, action, out=out, step_ctr=step_ctr)
            return next_state, obs, reward, flags, self.belief_histogram(next_state, all_reduce=all_reduce)
        bins = self._hist_n_bins()
        if hist_out is None:
            hist = self._empty((bins,), torch.int64)
        else:
            hist = hist_out
            if hist.dtype != torch.int64 or hist.numel() != bins or not hist.is_contiguous() or hist.device != self.device:
                raise ValueError("hist_out must be a contiguous int64[%d] tensor on %s" % (bins, self.device))
        # the sink structs are built once (per stream for the local one): an eager caller of a 2^20-env step is host-bound
        if all_reduce == "fused":
            sink = self.__dict__.get("_sink_fused")
            if sink is None:
                st = self._fused_hist_state()
                sink = self._sink_fused = _lib.HistSink()
                sink.scratch, sink.d_peer_bufs = _lib.ptr(st["scratch"]), st["hdl"].buffer_ptrs_dev
                sink.world, sink.rank, sink.wait = st["hdl"].world_size, st["hdl"].rank, 1
        else:
            stream = self._stream()
            sinks = self.__dict__.setdefault("_sinks_local", {})
            sink = sinks.get(stream)
            if sink is None:
                sink = sinks[stream] = _lib.HistSink()
                sink.scratch = _lib.ptr(self._local_hist_scratch())
        sink.hist_out = hist.data_ptr()
        ctr = self._next_ctr() if step_ctr is None else int(step_ctr)
        with self._guard():
            self._c_step_hist(state, action, next_state, obs, reward, flags, n, ctr, sink)
        if all_reduce is True:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        return next_state, obs, reward, flags, hist

    def _hist_n_bins(self):
        bins = self.__dict__.get("_hist_bins_cached")
        if bins is None:
            p0, p1 = self._hist_args()
            bins = self._hist_bins_cached = _lib.lib().pomdp_belief_hist_bins(self.kind, p0, p1)
        return bins

    def _local_hist_scratch(self):
        """The self-cleaning scratch of the one-launch histogram calls, one per stream (calls on one stream are ordered)."""
        stream = self._stream()
        scratches = self.__dict__.setdefault("_hist_scratch", {})
        scratch = scratches.get(stream)
        if scratch is None:
            scratch = scratches[stream] = torch.zeros(self._FUSED_HIST_BINS + 2, dtype=torch.int64, device=self.device)
        return scratch

    def init_states(self, n=None, out=None, mask=None, step_ctr=None):
        """Batched ``_get_init_state``: fresh initial states (and the reset observation)."""
        n = self.batch_size if n is None else int(n)
        if out is None:
            shape = (n, self.state_words) if self.state_words > 1 else (n,)
            out = (self._empty(shape, torch.int32), self._empty((n,), torch.int32))
        state, obs = out
        ctr = self._next_ctr() if step_ctr is None else int(step_ctr)
        with self._guard():
            self._c_reset(state, obs, mask, n, ctr)
        return state, obs

    def sample_legal_actions(self, state=None, out=None, step_ctr=None):
        """Batched ``np.random.choice(env._generate_legal())``: one uniformly drawn legal action per
        particle (the reference's list order decides which), as int32[n].  Uses the POLICY draw of
        ``step_ctr``; the default is the counter the NEXT ``simulate``/``step`` call will use, so
        ``a = sample_legal_actions(s); simulate(s, a)`` is one rollout step."""
        state = self.state if state is None else state
        n = state.shape[0]
        action = self._empty((n,), torch.int32) if out is None else out
        ctr = ((self._step_ctr + 1) & 0xFFFFFFFF) if step_ctr is None else int(step_ctr)
        with self._guard():
            self._c_policy(state, action, n, ctr)
        return action

    def rollout(self, state=None, max_steps=100, discount=None, out=None, step_ctr=None, first_action=None, policy="legal",
                **policy_kw):
        """Monte-Carlo rollouts under the uniform-legal policy, fused into ONE kernel (states stay
        in registers): until done or ``max_steps``,
        ``a = choice(_generate_legal()); ob, rw, done = step(a); ret += rw * disc; disc *= discount``.

        Returns (final_state, ret float64[n], steps int32[n], flags int32[n]).  ``discount`` defaults
        to the env's ``_discount``.  Draw for draw identical to ``max_steps`` rounds of
        ``sample_legal_actions`` + ``simulate`` with counters step_ctr, step_ctr + 1, ...
        ``first_action`` (int32[n]): step 0 takes these actions instead of a policy draw, so ``ret`` is a
        Monte-Carlo sample of Q(s, a) -- POMCP's "simulate a, then roll out" in one launch.
        ``policy="preferred"``: actions are drawn from ``_generate_preferred(history)`` instead (RockSample with
        ``use_heuristic``, Tag; the other envs' preferred set is the legal set); see ``_c_rollout_preferred`` of the
        env for the extra keywords."""
        if policy not in ("legal", "preferred"):
            raise ValueError("policy must be 'legal' or 'preferred'")
        state = self.state if state is None else state
        n = state.shape[0]
        if out is None:
            out = (torch.empty_like(state), self._empty((n,), torch.float64), self._empty((n,), torch.int32),
                   self._empty((n,), torch.int32))
        final_state, ret, steps, flags = out
        if step_ctr is None:
            ctr = (self._step_ctr + 1) & 0xFFFFFFFF
            self._step_ctr = (self._step_ctr + int(max_steps)) & 0xFFFFFFFF
        else:
            ctr = int(step_ctr)
        if first_action is not None:
            first_action = torch.as_tensor(first_action, device=state.device).to(torch.int32).expand(n).contiguous()
        with self._guard():
            if policy == "preferred" and self._has_preferred_kernel():
                self._c_rollout_preferred(state, final_state, ret, steps, flags, n, ctr, max_steps,
                                          self._discount if discount is None else discount, first_action, **policy_kw)
            else:
                if policy_kw:
                    raise TypeError("unexpected keywords for the uniform-legal rollout: %s" % sorted(policy_kw))
                self._c_rollout(state, final_state, ret, steps, flags, n, ctr, max_steps,
                                self._discount if discount is None else discount, first_action)
        return final_state, ret, steps, flags

    def _has_preferred_kernel(self):
        """True where ``_generate_preferred`` differs from ``_generate_legal`` (RockSample with use_heuristic, Tag)."""
        return False

    def simulate_host(self, state, action, out, step_ctr=None, chunk=None, packed=False, n_streams=3, zero_copy=False,
                      pipeline=None):
        """G(s, a) on HOST buffers (pinned CPU tensors): what a numpy-holding caller of the
        reference does.  The batch is cut into chunks that are copied in, stepped and copied
        out on three rotating CUDA streams, so the H2D copy, the kernel and the D2H copy of
        neighbouring chunks overlap (PCIe is full duplex).  ``out`` = (next_state, obs,
        reward, flags) pinned CPU tensors -- or, with ``packed=True``, (next_state, result):
        8 instead of 16 bytes per env come back over PCIe (``unpack_result`` decodes).
        Returns after all results have landed.

        ``pipeline``: who drives the chunks.  ``"c"`` -- ONE C-ABI call, ``pomdp_step_packed_host``
        (include/pomdp_b200.h): the library owns the staging buffers and streams and issues every copy and
        launch itself (no interpreter time per chunk).  ``"python"`` -- this method issues the copies and
        launches through torch.  Default: ``"c"`` for packed results on a CUDA device, ``"python"`` otherwise
        (the four-array result has no host entry point).  Default chunk: 2^20 envs -- measured on the B200
        box, smaller chunks lose more to the fixed cost of each copy than they gain in pipeline fill and
        drain (scripts/exp_host_pipe.py, scripts/exp_pcie_chunks.py), and the link carries ~52 GB/s one
        way but only ~75 GB/s both ways together, which is what bounds this call.

        ``zero_copy=True``: no staging at all -- ONE kernel launch whose loads and stores go straight
        to the pinned host buffers over PCIe (pinned memory is mapped into the device address space
        under unified addressing), so reading the inputs and writing the results overlap at the
        granularity of a warp instead of a chunk (measured 7 % faster than the copy pipeline for
        packed RockSample(11,11) results; hybrids -- copy engine one way, kernel the other -- were
        slower than both)."""
        n = action.shape[0]
        ctr = self._next_ctr() if step_ctr is None else int(step_ctr)
        if zero_copy:
            for t in (state, action) + tuple(out[:2 if packed else 4]):
                if not t.is_pinned():
                    raise ValueError("zero_copy needs pinned host tensors")
            with self._guard():
                if packed:
                    self._c_step_packed(state, action, out[0], out[1], n, ctr)
                else:
                    self._c_step(state, action, out[0], out[1], out[2], out[3], n, ctr)
                torch.cuda.current_stream(self.device).synchronize()
            return out
        if pipeline is None:
            pipeline = "c" if (packed and self.device.type == "cuda" and self.kind != _lib.KIND_BATTLESHIP) else "python"
        if pipeline == "c":
            if not packed:
                raise ValueError("the C host pipeline returns packed results (packed=True)")
            for t in (state, action, out[0], out[1]):
                if t.device.type != "cpu" or not t.is_contiguous() or t.dtype != torch.int32:
                    raise ValueError("simulate_host(pipeline='c') needs contiguous int32 CPU tensors")
            pipe = self._host_pipe(chunk or (1 << 20), n_streams)
            with self._guard():
                _lib.check(_lib.lib().pomdp_step_packed_host(
                    pipe, self.kind, ctypes.addressof(self._params), _lib.ptr(getattr(self, "_table", None)),
                    state.data_ptr(), action.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), n, self.global_offset,
                    self._seed, ctr), "pomdp_step_packed_host")
            return out
        if pipeline != "python":
            raise ValueError("pipeline must be 'c' or 'python'")
        ws = self._host_ws(min(chunk or (1 << 20), max(n, 1)), n_streams)
        n_out = 2 if packed else 4
        base_off = self.global_offset
        with self._guard():
            cur = torch.cuda.current_stream(self.device)
            for st in ws["streams"]:
                st.wait_stream(cur)
            for ci, lo in enumerate(range(0, n, ws["chunk"])):
                hi = min(n, lo + ws["chunk"])
                m = hi - lo
                slot = ws["slots"][ci % len(ws["slots"])]
                with torch.cuda.stream(ws["streams"][ci % len(ws["streams"])]):
                    d = [b[:m] for b in slot]
                    self.global_offset = base_off + lo
                    d[0].copy_(state[lo:hi], non_blocking=True)
                    d[1].copy_(action[lo:hi], non_blocking=True)
                    if packed:
                        self._c_step_packed(d[0], d[1], d[2], d[3], m, ctr)
                    else:
                        self._c_step(d[0], d[1], d[2], d[3], d[4], d[5], m, ctr)
                    for k in range(n_out):
                        out[k][lo:hi].copy_(d[2 + k], non_blocking=True)
            self.global_offset = base_off
            for st in ws["streams"]:
                cur.wait_stream(st)
            cur.synchronize()
        return out

    def _host_pipe(self, chunk, n_slots=3):
        """The C library's staging pipe for ``simulate_host(pipeline='c')``, created once per (chunk, n_slots)."""
        key = (int(chunk), int(n_slots))
        cur = getattr(self, "_hpipe", None)
        if cur is not None and cur[0] == key:
            return cur[1]
        self._close_host_pipe()
        h = ctypes.c_void_p()
        with self._guard():
            _lib.check(_lib.lib().pomdp_host_pipe_create(self.state_words, key[0], key[1], ctypes.byref(h)), "pomdp_host_pipe_create")
        self._hpipe = (key, h)
        return h

    def _close_host_pipe(self):
        cur = getattr(self, "_hpipe", None)
        if cur is not None:
            self._hpipe = None
            try:
                _lib.lib().pomdp_host_pipe_destroy(cur[1])
            except Exception:       # interpreter shutdown
                pass

    def __del__(self):
        self._close_host_pipe()

    def _host_ws(self, chunk, n_streams=3):
        ws = getattr(self, "_hws", None)
        if ws is None or ws["chunk"] != chunk or len(ws["streams"]) != n_streams:
            sshape = (chunk, self.state_words) if self.state_words > 1 else (chunk,)
            slots = [(self._empty(sshape, torch.int32), self._empty((chunk,), torch.int32),
                      self._empty(sshape, torch.int32), self._empty((chunk,), torch.int32),
                      self._empty((chunk,), torch.float32), self._empty((chunk,), torch.int32)) for _ in range(n_streams)]
            ws = self._hws = {"chunk": chunk, "slots": slots,
                              "streams": [torch.cuda.Stream(self.device) for _ in range(n_streams)]}
        return ws

    # ------------------------------------------------------------------ gym surface ---
    def seed(self, seed=None):
        """The reference seeds numpy's global RNG (e.g. rock.py:120-121); here the seed keys
        the stateless Philox stream and restarts the step counter."""
        self._seed = _fresh_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        self._step_ctr = 0
        return [seed]

    # scalar (drop-in) mode keeps its one instance in a small PINNED HOST buffer that the kernels read and write
    # directly (zero-copy over PCIe): [state(W) | next_state(W) | action | obs | reward | flags] as int32.  A step is
    # then one launch + one stream synchronize, and every conversion to the reference's state formats is plain
    # Python on host integers -- no device tensors, no extra copies.
    def _ensure_io(self):
        if getattr(self, "_io", None) is not None:
            return
        W = self.state_words
        io = torch.zeros(2 * W + 4, dtype=torch.int32)
        if self.device.type == "cuda":
            io = io.pin_memory()
        self._io, self._io_np, self._io_f = io, io.numpy(), io.view(torch.float32)
        shape = (1, W) if W > 1 else (1,)
        self._io_state, self._io_next = io[:W].view(shape), io[W:2 * W].view(shape)
        self._io_act, self._io_obs = io[2 * W:2 * W + 1], io[2 * W + 1:2 * W + 2]
        self._io_rw, self._io_fl = self._io_f[2 * W + 2:2 * W + 3], io[2 * W + 3:2 * W + 4]

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def _host_words(self):
        """the scalar instance's packed state as unsigned Python ints"""
        return [int(w) & 0xFFFFFFFF for w in self._io_np[:self.state_words]]

    def reset(self, mask=None, seed=None, options=None):
        """``reset() -> ob`` (the reference's protocol).  ``seed`` / ``options`` are accepted for callers written against
        newer gym versions: a seed is applied through ``seed()``; options may carry ``{"mask": ...}``."""
        if seed is not None:
            self.seed(seed)
        if mask is None and isinstance(options, dict):
            mask = options.get("mask")
        if self._scalar:
            self._ensure_io()
            with self._guard():
                self._c_reset(self._io_state, self._io_obs, None, 1, self._next_ctr())
                self._sync()
            self.state, self.flags = self._io_state, self._io_fl
            self._io_np[2 * self.state_words + 3] = 0
            self._on_reset()
            self.done = False
            return int(self._io_np[2 * self.state_words + 1])
        if self.state is None or mask is None:
            self.state, obs = self.init_states(self.batch_size)
            self.flags = torch.zeros(self.batch_size, dtype=torch.int32, device=self.device)
        else:
            m = mask.to(device=self.device, dtype=torch.uint8)
            obs = torch.zeros(self.batch_size, dtype=torch.int32, device=self.device)
            self.init_states(self.batch_size, out=(self.state, obs), mask=m)
            self.flags = torch.where(m.bool(), torch.zeros_like(self.flags), self.flags)
        self._on_reset()
        return obs

    def _on_reset(self):
        pass

    def step(self, action):
        if self._scalar:
            return self._step_scalar(action)
        if self.state is None:
            raise AssertionError("step() before reset()")
        action = torch.as_tensor(action, device=self.device).to(torch.int32).contiguous()
        if action.shape != (self.batch_size,):
            raise ValueError("action must have shape (%d,), got %s" % (self.batch_size, tuple(action.shape)))
        next_state, obs, reward, flags = self.simulate(self.state, action)
        self.state, self.flags = next_state, flags
        done = (flags & _lib.FLAG_DONE) != 0
        return obs, reward, done, {"state": next_state, "flags": flags}

    def _step_scalar(self, action):
        # the reference's two asserts (rock.py:125-126 and siblings)
        assert self.action_space.contains(action)
        assert self.done is False
        if self.state is None:
            raise AttributeError("%s has no state: call reset() first" % type(self).__name__)
        W, io = self.state_words, self._io_np
        io[2 * W] = int(action)
        with self._guard():
            self._c_step(self._io_state, self._io_act, self._io_next, self._io_obs, self._io_rw, self._io_fl, 1, self._next_ctr())
            self._sync()
        fl = int(io[2 * W + 3])
        self._raise_for_flags(fl)                 # the state is still the pre-step one if this raises
        io[:W] = io[W:2 * W]
        self.last_action = int(action)
        self.done = bool(fl & _lib.FLAG_DONE)
        ob = int(io[2 * W + 1])
        self._after_scalar_step(int(action), ob)
        return ob, self._reward_to_py(float(self._io_f[2 * W + 2]), int(action)), self.done, {"state": self._info_state()}

    def _after_scalar_step(self, action, ob):
        pass

    def _info_state(self):
        return self._state_to_ref(self._host_words())

    def _set_state(self, state):
        """Batched: packed int32 tensor (copied).  Scalar: the reference's own state format."""
        if self._scalar:
            self.done = False
            self._ensure_io()
            self._io_state.copy_(self._state_from_ref(state).reshape(self._io_state.shape))
            self.state, self.flags = self._io_state, self._io_fl
            self._io_np[2 * self.state_words + 3] = 0
            return
        else:
            state = torch.as_tensor(state, device=self.device).to(torch.int32)
            expect = (self.batch_size, self.state_words) if self.state_words > 1 else (self.batch_size,)
            if tuple(state.shape) != expect:
                raise ValueError("state must have shape %s, got %s" % (expect, tuple(state.shape)))
            self.state = state.clone().contiguous()
        self.flags = torch.zeros(self.batch_size, dtype=torch.int32, device=self.device)

    def _get_init_state(self):
        state, _ = self.init_states(self.batch_size)
        if self._scalar:
            return self._state_to_ref([int(w) & 0xFFFFFFFF for w in state.reshape(-1).tolist()])
        return state

    def render(self, mode="ansi", close=False):
        if close:
            return
        if self._scalar and self.state is not None:
            print(type(self).__name__, self._info_state())

    def close(self):
        self._close_host_pipe()

    # ------------------------------------------------------------- belief histogram ---
    def _hist_args(self):
        raise NotImplementedError

    def belief_histogram(self, state=None, all_reduce=False):
        """int64 counts over the particle set (bins: include/pomdp_b200.h).  ``all_reduce=True``: the counts are summed
        over all ranks of the default process group by NCCL -- the only collective on the path.  ``all_reduce="fused"``:
        the same sum, but made INSIDE the histogram kernel over NVLink / NVSwitch peer memory
        (``pomdp_belief_hist_allreduce``: the last CTA of every rank adds the rank's counts into every rank's buffer,
        signals its arrival to every peer and waits for theirs; torch symmetric memory provides the peer mappings) --
        ONE kernel, no collective launch."""
        state = self.state if state is None else state
        if all_reduce == "fused":
            return self._belief_histogram_fused(state)
        p0, p1 = self._hist_args()
        L = _lib.lib()
        bins = self._hist_n_bins()
        # ONE launch (pomdp_belief_hist_once): the last arrival at every bin moves its count from a self-cleaning scratch to
        # the result, so no zero-fill kernel runs before it.  One scratch per stream: calls on one stream are ordered.
        scratch = self._local_hist_scratch()
        hist = torch.empty(bins, dtype=torch.int64, device=self.device)
        n = state.shape[0]
        with self._guard():
            _lib.check(L.pomdp_belief_hist_once(self.kind, p0, p1, _lib.ptr(state), self.state_words, n, _lib.ptr(scratch),
                                                _lib.ptr(hist), self._stream()), "pomdp_belief_hist_once")
        if all_reduce:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        return hist

    _FUSED_HIST_BINS = 512          # POMDP_HIST_MAX_BINS
    _FUSED_HIST_RANKS = 64          # POMDP_HIST_MAX_RANKS

    def _fused_hist_state(self):
        """One symmetric buffer per rank (two alternating result slots + a row of arrival counters: POMDP_HIST_SYMM_WORDS),
        its peer table and this rank's scratch, created once.  A COLLECTIVE call: every rank of the default group must
        reach it."""
        st = getattr(self, "_fhist", None)
        if st is None:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("belief_histogram(all_reduce='fused') needs an initialised torch.distributed process group")
            buf = symm_mem.empty((2 * self._FUSED_HIST_BINS + self._FUSED_HIST_RANKS,), dtype=torch.int64, device=self.device)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            if hdl.world_size > self._FUSED_HIST_RANKS:
                raise RuntimeError("belief_histogram(all_reduce='fused') supports at most %d ranks" % self._FUSED_HIST_RANKS)
            buf.zero_()
            hdl.barrier(channel=0)          # every rank's slots and counters are zero before any rank touches them
            scratch = torch.zeros(self._FUSED_HIST_BINS + 2, dtype=torch.int64, device=self.device)
            st = self._fhist = {"buf": buf, "hdl": hdl, "scratch": scratch}
        return st

    def _belief_histogram_fused(self, state, out=None):
        """ONE launch: the kernel counts, adds the counts into every rank's buffer over NVLink, signals and waits for its
        peers and writes the global counts to ``out``.  Nothing about the launch changes from call to call (the epoch is
        kept in device memory), so it can be captured in a CUDA graph."""
        p0, p1 = self._hist_args()
        L = _lib.lib()
        bins = L.pomdp_belief_hist_bins(self.kind, p0, p1)
        st = self._fused_hist_state()
        hdl = st["hdl"]
        if out is None:
            out = torch.empty(bins, dtype=torch.int64, device=self.device)
        n = state.shape[0]
        with self._guard():
            _lib.check(L.pomdp_belief_hist_allreduce(
                self.kind, p0, p1, _lib.ptr(state), self.state_words, n, _lib.ptr(st["scratch"]), hdl.buffer_ptrs_dev,
                hdl.world_size, hdl.rank, 1, _lib.ptr(out), self._stream()), "pomdp_belief_hist_allreduce")
        return out


class _NullCtx(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
