"""RockSample on the GPU: host side of ``pomdp_rock_step`` / ``pomdp_rock_reset``.

Stands in for gym_pomdp/envs/rock.py ``RockEnv`` (96-407) and ``StochasticRockEnv``
(428-504).  Packed state (include/pomdp_b200.h): bits 0-3 x, 4-7 y, two bits per rock
(0b11 bad, 0b00 collected, 0b01 good), top bit done; one int32 word for <= 11 rocks, two
otherwise.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..geometry import Coord, Grid
from ..spaces import Discrete
from .base import BatchedPomdpEnv

NULL, BAD, GOOD = 0, 1, 2          # rock.py:12-15
SAMPLE = 4                         # rock.py:18-23


class RockBeliefStats(object):
    """Per-rock belief side-statistics of a batch (rock.py:78-86): ``count``, ``measured`` int32[n, k];
    ``lkv``, ``lkw``, ``prob_valuable`` float64[n, k].  Fresh values as in ``Rock.__init__``."""

    def __init__(self, n, k, device, pinned_host=False):
        """pinned_host: page-locked HOST tensors the kernels update in place over PCIe (the single-instance mode keeps its
        one env's planes there, next to its pinned state buffer)."""
        kw = dict(device="cpu", pin_memory=True) if pinned_host else dict(device=device)
        self.count = torch.zeros((n, k), dtype=torch.int32, **kw)
        self.measured = torch.zeros((n, k), dtype=torch.int32, **kw)
        self.lkv = torch.ones((n, k), dtype=torch.float64, **kw)
        self.lkw = torch.ones((n, k), dtype=torch.float64, **kw)
        self.prob_valuable = torch.full((n, k), .5, dtype=torch.float64, **kw)

    def reset(self, mask=None):
        """Back to the fresh values, for all envs or those selected by ``mask`` (bool[n])."""
        sel = slice(None) if mask is None else mask.to(self.count.device).bool()
        self.count[sel] = 0
        self.measured[sel] = 0
        self.lkv[sel] = 1.
        self.lkw[sel] = 1.
        self.prob_valuable[sel] = .5

    def arrays(self):
        return self.count, self.measured, self.lkv, self.lkw, self.prob_valuable


class RockHistory(object):
    """What ``_generate_preferred`` needs from the caller's History, for a whole batch: per rock the two running totals
    of rock.py:302-309 / 325-331 over the transitions that checked it (``check_totals`` int32[n, k]: low 16 bits
    #GOOD - #BAD of ``next_observation``, high 16 bits #(next_observation == GOOD) - #(next_observation != GOOD and
    observation == BAD)) and the observation the next transition will record as ``observation`` (``prev_obs`` int32[n])."""

    def __init__(self, n, k, device):
        self.check_totals = torch.zeros((n, k), dtype=torch.int32, device=device)
        self.prev_obs = torch.zeros(n, dtype=torch.int32, device=device)          # RockEnv.reset returns Obs.NULL

    def reset(self, mask=None):
        sel = slice(None) if mask is None else mask.to(self.prev_obs.device).bool()
        self.check_totals[sel] = 0
        self.prev_obs[sel] = 0

    def totals(self):
        """(tot_sample, tot_dir) int32[n, k]"""
        t = self.check_totals
        return ((t << 16) >> 16), (t >> 16)


class Transition(tuple):
    """rock.py:525-530: NamedTuple(observation, action, reward, next_observation, done)"""
    __slots__ = ()
    _fields = ("observation", "action", "reward", "next_observation", "done")

    def __new__(cls, observation, action, reward, next_observation, done):
        return tuple.__new__(cls, (observation, action, reward, next_observation, done))

    observation = property(lambda self: self[0])
    action = property(lambda self: self[1])
    reward = property(lambda self: self[2])
    next_observation = property(lambda self: self[3])
    done = property(lambda self: self[4])


class History(object):
    """rock.py:533-551"""

    def __init__(self, max_size=None):
        self._max_size = max_size
        self._history = []

    def __getitem__(self, item):
        return self._history[item]

    def append(self, transition):
        if self._max_size is not None and self.size > self._max_size:
            self._history.pop(0)
        self._history.append(transition)

    @property
    def size(self):
        return len(self._history)

    def __repr__(self):
        return "size:%d" % self.size


class RockEnv(BatchedPomdpEnv):
    kind = _lib.KIND_ROCK
    _abi = "rock"
    _stochastic = False

    def __init__(self, board_size=7, num_rocks=8, use_heuristic=False, batch_size=None, device="cuda", seed=None,
                 global_offset=0, p_move=0.8, track_belief_stats=False, track_history=False, history_next_is_reward=False):
        super().__init__(batch_size, device, seed, global_offset)
        # batched mode: keep rock.py's per-rock side-stats current; track_history also keeps the per-rock totals
        # _generate_preferred reads from the History (as if every step appended Transition(ob, a, rw, next_ob, done); with
        # history_next_is_reward the reference's own positional Transition(ob, a, next_ob, rw, done) of rock.py:566)
        self.track_history = bool(track_history)
        self.track_belief_stats = bool(track_belief_stats) or self.track_history
        self.history_next_is_reward = bool(history_next_is_reward)
        self.belief_stats = None
        self.history = None
        self.num_rocks = num_rocks
        self._use_heuristic = use_heuristic
        self._params = _lib.RockParams(board_size, num_rocks, int(self._stochastic), 0, float(p_move))
        L = _lib.lib()
        words = L.pomdp_rock_state_words(ctypes.byref(self._params))
        # rock.py:101 -- the reference asserts on an unknown configuration
        assert words > 0, L.pomdp_last_error().decode()
        self.state_words = words
        nbytes = L.pomdp_rock_table_bytes(ctypes.byref(self._params))
        host = np.zeros(nbytes, dtype=np.uint8)
        _lib.check(L.pomdp_rock_build_table(ctypes.byref(self._params), host.ctypes.data), "pomdp_rock_build_table")
        self._table_host = host
        self._table = torch.from_numpy(host.copy()).to(self.device)   # header + LUT, staged to smem by TMA per CTA
        self._grid_map = host[:256].view(np.int8)                     # [x | y << 4] -> rock id
        self._rock_pos = [Coord(int(b) & 15, int(b) >> 4) for b in host[256:256 + num_rocks]]
        thr_m1 = host[272:400].view(np.uint32)
        self._eff_T = thr_m1.astype(np.int64) + 1                     # ceil(eff(d) * 2^32)
        self.grid = Grid(board_size, board_size)
        for x in range(board_size):                                   # rock.py:110-111: rock ids in the board (later ids win)
            for y in range(board_size):
                self.grid.board[x, y] = self._grid_map[x | (y << 4)]
        self.action_space = Discrete(5 + num_rocks)                   # rock.py:113
        self.observation_space = Discrete(3)                          # rock.py:114
        self._discount = .95
        self._reward_range = 20
        self._penalization = 0 if self._stochastic else -100
        self._query = 0
        self._sstats = None   # scalar mode: per-rock belief side-stats (rock.py:82-86) as pinned host planes the kernel updates
        self._legal_act = host[400:432].copy()                        # list position of _generate_legal -> action id

    # -------------------------------------------------------------------- C calls ---
    def _c_head(self):
        return (ctypes.byref(self._params), _lib.ptr(self._table))

    _c_query_head = _c_head

    def _c_step_hist(self, state, action, next_state, obs, reward, flags, n, ctr, sink):
        _lib.check(_lib.lib().pomdp_rock_step_hist(
            ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state),
            _lib.ptr(obs), _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, ctypes.byref(sink),
            self._stream()), "pomdp_rock_step_hist")

    def _c_step(self, state, action, next_state, obs, reward, flags, n, ctr):
        _lib.check(_lib.lib().pomdp_rock_step(
            ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state),
            _lib.ptr(obs), _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, self._stream()),
            "pomdp_rock_step")
        if self._scalar and self._sstats is not None and next_state is self._io_next:
            # the single instance's belief side-statistics (rock.py:177-191): the same kernel as in batched mode, queued
            # behind the step on the same stream, on the pinned planes
            st = self._sstats
            _lib.check(_lib.lib().pomdp_rock_belief_update(
                ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(next_state), _lib.ptr(action), _lib.ptr(obs),
                _lib.ptr(st.count), _lib.ptr(st.measured), _lib.ptr(st.lkv), _lib.ptr(st.lkw), _lib.ptr(st.prob_valuable), 1,
                self._stream()), "pomdp_rock_belief_update")

    def _c_reset(self, state, obs, mask, n, ctr):
        _lib.check(_lib.lib().pomdp_rock_reset(
            ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state), _lib.ptr(obs), _lib.ptr(mask), n,
            self.global_offset, self._seed, ctr, self._stream()), "pomdp_rock_reset")

    def _hist_args(self):
        return self.num_rocks, self.state_words

    # ------------------------------------------------- belief side-statistics (batched) ---
    def new_belief_stats(self, n=None):
        return RockBeliefStats(self.batch_size if n is None else int(n), self.num_rocks, self.device)

    def update_belief_stats(self, stats, next_state, action, obs):
        """rock.py:177-191 for a whole batch: every env whose ``action`` was a check that produced a reading updates
        that rock's measured / count / lkv / lkw / prob_valuable in ``stats`` (in place; one kernel)."""
        n = next_state.shape[0]
        action = torch.as_tensor(action, device=self.device).to(torch.int32).contiguous()
        obs = torch.as_tensor(obs, device=self.device).to(torch.int32).contiguous()
        if action.shape != (n,) or obs.shape != (n,):
            raise ValueError("action and obs must have shape (%d,), got %s and %s" % (n, tuple(action.shape), tuple(obs.shape)))
        with self._guard():
            _lib.check(_lib.lib().pomdp_rock_belief_update(
                ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(next_state), _lib.ptr(action), _lib.ptr(obs),
                _lib.ptr(stats.count), _lib.ptr(stats.measured), _lib.ptr(stats.lkv), _lib.ptr(stats.lkw),
                _lib.ptr(stats.prob_valuable), n, self._stream()), "pomdp_rock_belief_update")
        return stats

    # ------------------------------------------------------ heuristic action sets (batched) ---
    def new_history(self, n=None):
        return RockHistory(self.batch_size if n is None else int(n), self.num_rocks, self.device)

    def update_history(self, history, action, next_obs_field, obs_field=None):
        """One transition appended to every env's history: the checked rock's totals (rock.py:302-309, 325-331) from the
        two fields as the caller's Transition holds them -- ``obs_field`` defaults to ``history.prev_obs``,
        ``next_obs_field`` is the next observation (or the reward, for the reference's positional Transition) -- then
        ``prev_obs`` advances to the step's observation if ``next_obs_field`` is that (pass ``obs_field`` explicitly
        otherwise and set ``history.prev_obs`` yourself)."""
        n = history.prev_obs.shape[0]
        action = torch.as_tensor(action, device=self.device).to(torch.int32).contiguous()
        nf = torch.as_tensor(next_obs_field, device=self.device).to(torch.int32).contiguous()
        of = history.prev_obs if obs_field is None else torch.as_tensor(obs_field, device=self.device).to(torch.int32).contiguous()
        with self._guard():
            _lib.check(_lib.lib().pomdp_rock_history_update(
                ctypes.byref(self._params), _lib.ptr(of), _lib.ptr(action), _lib.ptr(nf), _lib.ptr(history.check_totals), n,
                self._stream()), "pomdp_rock_history_update")
        return history

    def _preferred_call(self, fn_name, state, stats, history, out, *tail):
        n = state.shape[0]
        fn = getattr(_lib.lib(), fn_name)
        with self._guard():
            _lib.check(fn(ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state.contiguous()),
                          _lib.ptr(None if stats is None else stats.count), _lib.ptr(None if stats is None else stats.measured),
                          _lib.ptr(None if stats is None else stats.prob_valuable),
                          _lib.ptr(None if history is None else history.check_totals), _lib.ptr(out), n, *tail, self._stream()),
                       fn_name)
        return out

    def preferred_mask_words(self, state=None, stats=None, history=None):
        """``_generate_preferred(history)`` (rock.py:293-374, use_heuristic=True) for every particle as one bit mask over
        action ids, int32[n]; 0 = the preferred list is empty and the reference returns ``_generate_legal()``.  ``stats``
        (RockBeliefStats) / ``history`` (RockHistory) default to the env's own planes, None = fresh / empty."""
        state = self.state if state is None else state
        stats = self.belief_stats if stats is None else stats
        history = self.history if history is None else history
        return self._preferred_call("pomdp_rock_preferred_mask", state, stats, history,
                                    self._empty((state.shape[0],), torch.int32))

    def sample_preferred_actions(self, state=None, stats=None, history=None, out=None, step_ctr=None):
        """Batched ``np.random.choice(env._generate_preferred(history))``: int32[n], the POLICY draw of ``step_ctr`` (default:
        the counter the next ``simulate``/``step`` call uses, like ``sample_legal_actions``)."""
        state = self.state if state is None else state
        stats = self.belief_stats if stats is None else stats
        history = self.history if history is None else history
        action = self._empty((state.shape[0],), torch.int32) if out is None else out
        ctr = ((self._step_ctr + 1) & 0xFFFFFFFF) if step_ctr is None else int(step_ctr)
        return self._preferred_call("pomdp_rock_policy_preferred", state, stats, history, action, self.global_offset,
                                    self._seed, ctr)

    def _has_preferred_kernel(self):
        return bool(self._use_heuristic)

    def _c_rollout_preferred(self, state, final_state, ret, steps, flags, n, ctr, max_steps, discount, first_action=None,
                             stats=None, history=None, next_is_reward=None):
        """rock.py:557-572 with use_heuristic=True, fused: ``stats`` / ``history`` planes (read at the start, updated in
        place; None = fresh Rock.__init__ values / an empty history); ``next_is_reward`` selects what the transitions'
        ``next_observation`` field holds (default: the env's ``history_next_is_reward``)."""
        planes = _lib.RockHeuristicPlanes()
        if stats is not None:
            planes.count, planes.measured = stats.count.data_ptr(), stats.measured.data_ptr()
            planes.lkv, planes.lkw, planes.prob_valuable = stats.lkv.data_ptr(), stats.lkw.data_ptr(), stats.prob_valuable.data_ptr()
        if history is not None:
            planes.check_totals, planes.prev_obs = history.check_totals.data_ptr(), history.prev_obs.data_ptr()
        if stats is None and history is None:
            # fresh planes in, none out: the kernel keeps its per-rock side-state as 32-byte records in this scratch
            need = n * self.num_rocks * 32
            buf = getattr(self, "_heur_scratch", None)
            if buf is None or buf.numel() < need or buf.device != state.device:
                buf = self._heur_scratch = torch.empty(max(need, 32), dtype=torch.uint8, device=state.device)
            planes.scratch = buf.data_ptr()
        nir = self.history_next_is_reward if next_is_reward is None else bool(next_is_reward)
        _lib.check(_lib.lib().pomdp_rock_rollout_preferred(
            ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state), _lib.ptr(first_action), ctypes.byref(planes),
            _lib.ptr(final_state), _lib.ptr(ret), _lib.ptr(steps), _lib.ptr(flags), n, self.global_offset, self._seed, ctr,
            int(max_steps), float(discount), int(nir), self._stream()), "pomdp_rock_rollout_preferred")

    # ---------------------------------------------------------------------- codec ---
    def pack(self, x, y, status, done=None):
        """x, y int[n]; status int[n, k] in {-1, 0, +1} -> packed int32[n] / int32[n, 2]."""
        x = torch.as_tensor(x, device=self.device).to(torch.int64)
        y = torch.as_tensor(y, device=self.device).to(torch.int64)
        status = torch.as_tensor(status, device=self.device).to(torch.int64)
        v = x | (y << 4)
        for i in range(self.num_rocks):
            v = v | ((status[:, i] & 3) << (8 + 2 * i))
        if done is not None:
            top = 31 if self.state_words == 1 else 63
            v = v | (torch.as_tensor(done, device=self.device).to(torch.int64) << top)
        return self._words_from_int64(v)

    def _words_from_int64(self, v):
        if self.state_words == 1:
            return ((v + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)
        lo = ((v & 0xFFFFFFFF) + 2 ** 31) % 2 ** 32 - 2 ** 31
        hi = v >> 32
        return torch.stack([lo, hi], dim=1).to(torch.int32).contiguous()

    def _int64_from_words(self, words):
        words = words.to(torch.int64)
        if self.state_words == 1:
            return words & 0xFFFFFFFF
        return (words[:, 0] & 0xFFFFFFFF) | (words[:, 1] << 32)

    def unpack(self, words):
        """packed -> (x, y, status[n, k], done)"""
        v = self._int64_from_words(words)
        x, y = v & 15, (v >> 4) & 15
        codes = torch.stack([(v >> (8 + 2 * i)) & 3 for i in range(self.num_rocks)], dim=1)
        status = torch.where(codes == 3, torch.full_like(codes, -1), codes)
        top = 31 if self.state_words == 1 else 63
        return x.to(torch.int32), y.to(torch.int32), status.to(torch.int32), ((v >> top) & 1).bool()

    def to_array_form(self, words):
        """``[agent_idx, status...]`` rows as in rock.py:205-210 / 376-381."""
        x, y, status, _ = self.unpack(words)
        idx = self.grid.x_size * y + x
        return torch.cat([idx[:, None], status], dim=1)

    # ---------------------------------------------------------------- scalar mode ---
    def reset(self, mask=None, seed=None, options=None):
        if mask is None and isinstance(options, dict):
            mask = options.get("mask")
        obs = super().reset(mask, seed)
        if self.track_belief_stats and not self._scalar:
            if self.belief_stats is None:
                self.belief_stats = self.new_belief_stats()
            else:
                self.belief_stats.reset(mask)
        if self.track_history and not self._scalar:
            if self.history is None:
                self.history = self.new_history()
            else:
                self.history.reset(mask)
        return obs

    def step(self, action):
        if not self._scalar and (self.track_belief_stats or self.track_history):
            action = torch.as_tensor(action, device=self.device).to(torch.int32).contiguous()
        out = super().step(action)
        if self.track_belief_stats and not self._scalar:
            self.update_belief_stats(self.belief_stats, self.state, action, out[0])
            out[3]["belief_stats"] = self.belief_stats
        if self.track_history and not self._scalar:
            nf = out[1].to(torch.int32) if self.history_next_is_reward else out[0]
            self.update_history(self.history, action, nf)
            self.history.prev_obs.copy_(out[0])
            out[3]["history"] = self.history
        return out

    def _on_reset(self):
        self._query = 0
        self.last_action = SAMPLE
        if self._scalar:
            if self._sstats is None:
                self._sstats = RockBeliefStats(1, self.num_rocks, self.device, pinned_host=self.device.type == "cuda")
            else:
                self._sstats.reset()

    @property
    def _side(self):
        """the single instance's per-rock side-statistics as the reference's Rock attributes (rock.py:82-86)"""
        st = self._sstats
        if st is None:
            return None
        c, m = st.count[0].tolist(), st.measured[0].tolist()
        v, w, p = st.lkv[0].tolist(), st.lkw[0].tolist(), st.prob_valuable[0].tolist()
        return [dict(count=c[i], measured=m[i], lkw=w[i], lkv=v[i], prob_valuable=p[i]) for i in range(self.num_rocks)]

    def _decode_py(self, words):
        """host ints -> (x, y, [status...])"""
        v = words[0] | (words[1] << 32 if len(words) > 1 else 0)
        return v & 15, (v >> 4) & 15, [(0, 1, 0, -1)[(v >> (8 + 2 * i)) & 3] for i in range(self.num_rocks)]

    def _state_to_ref(self, words):
        """The reference's ``_encode_dict`` layout (rock.py:507-516)."""
        x, y, status = self._decode_py(words)
        side = self._side or [dict(count=0, measured=0, lkw=1., lkv=1., prob_valuable=.5)] * self.num_rocks
        rocks = [{"status": status[i], "pos": self._rock_pos[i], "count": side[i]["count"],
                  "measured": side[i]["measured"], "lkw": side[i]["lkw"], "lkv": side[i]["lkv"],
                  "prob_valuable": side[i]["prob_valuable"]} for i in range(self.num_rocks)]
        return {"agent_pos": Coord(x, y), "rocks": rocks, "target": -1}

    def _pack_ref(self, state):
        ax, ay = state["agent_pos"]
        return self.pack([ax], [ay], [[int(r["status"]) for r in state["rocks"]]])

    def _state_from_ref(self, state):
        if self._sstats is None:
            self._sstats = RockBeliefStats(1, self.num_rocks, self.device, pinned_host=self.device.type == "cuda")
        st, rocks = self._sstats, state["rocks"]
        st.count[0] = torch.tensor([int(r.get("count", 0)) for r in rocks], dtype=torch.int32)
        st.measured[0] = torch.tensor([int(r.get("measured", 0)) for r in rocks], dtype=torch.int32)
        st.lkw[0] = torch.tensor([float(r.get("lkw", 1.)) for r in rocks], dtype=torch.float64)
        st.lkv[0] = torch.tensor([float(r.get("lkv", 1.)) for r in rocks], dtype=torch.float64)
        st.prob_valuable[0] = torch.tensor([float(r.get("prob_valuable", .5)) for r in rocks], dtype=torch.float64)
        return self._pack_ref(state)

    def _reward_to_py(self, reward, action):
        return int(reward)

    def _raise_for_flags(self, flags):
        if flags & _lib.FLAG_BAD_STATE:
            # the reference dies with IndexError at rock.py:162 on a dangling grid id
            raise IndexError("list index out of range")

    @staticmethod
    def _efficiency(agent_pos, rock_pos, hed=20):
        """rock.py:383-387 (the kernels read eff(d) from the table built with this same formula on the host)"""
        d = Grid.euclidean_distance(agent_pos, rock_pos)
        return (1 + pow(2, -d / hed)) * .5

    def _after_scalar_step(self, action, ob):
        self._query += 1

    def _scalar_state_tensor(self):
        """the single instance's packed state where the kernels can read it"""
        return self._io_state

    # ------------------------------------------------------------- planner hooks ---
    def _generate_legal(self, state=None):
        """rock.py:273-291.  Scalar mode: the reference's list (same order).  Batched: bool
        mask [n, n_actions] (a set; the reference's duplicate entries for Rock(15,15)'s
        doubled rock collapse)."""
        if self._scalar and state is None:
            lst = torch.zeros(1, dtype=torch.int32, device=self.device)
            with self._guard():
                _lib.check(_lib.lib().pomdp_rock_legal_list(ctypes.byref(self._params), _lib.ptr(self._table),
                                                            _lib.ptr(self._scalar_state_tensor()), _lib.ptr(lst), 1, self._stream()),
                           "pomdp_rock_legal_list")
            word = int(lst[0]) & 0xFFFFFFFF
            return [int(self._legal_act[b]) for b in range(5 + self.num_rocks) if (word >> b) & 1]
        return self.legal_mask(state)

    def _generate_preferred(self, history):
        """rock.py:293-374.  Without ``use_heuristic``: ``_generate_legal()``.  Scalar mode: ``history`` is the caller's
        History of Transitions, read field by field exactly as the reference reads it (``action``, ``next_observation``
        and -- in the second loop's elif, rock.py:330 -- ``observation``); returns the reference's list.  Batched mode:
        ``history`` is a RockHistory (or None for the env's own / an empty one); returns bool[n, n_actions]."""
        if not self._use_heuristic:
            return self._generate_legal()
        if not self._scalar:
            words = self.preferred_mask_words(history=history).to(torch.int64) & 0xFFFFFFFF
            a = torch.arange(self.action_space.n, device=words.device)
            pref = ((words[:, None] >> a) & 1).bool()
            return torch.where((words == 0)[:, None], self.legal_mask(), pref)
        # the two per-rock totals the reference recomputes from the whole history on every call
        k = self.num_rocks
        ts, td = [0] * k, [0] * k
        for tr in history:
            r = tr.action - SAMPLE - 1
            if 0 <= r < k:
                if tr.next_observation == GOOD:
                    ts[r] += 1
                    td[r] += 1
                else:
                    if tr.next_observation == BAD:
                        ts[r] -= 1
                    if tr.observation == BAD:
                        td[r] -= 1
        hist = RockHistory(1, k, self.device)
        hist.check_totals.copy_(torch.tensor([[(s_ & 0xFFFF) | ((d_ & 0xFFFF) << 16) for s_, d_ in zip(ts, td)]],
                                             dtype=torch.int64).to(torch.int32))
        word = int(self.preferred_mask_words(self._scalar_state_tensor(), self._sstats, hist)[0]) & 0xFFFFFFFF
        if word == 0:
            return self._generate_legal()
        return [a for a in range(self.action_space.n) if (word >> a) & 1]

    @staticmethod
    def _select_target(rock_state, x_size):
        """rock.py:389-399: the nearest (2-norm: ``Grid.manhattan_distance`` is the Euclidean one) rock that is not
        collected and whose check count is not negative; -1 if none within 2 * x_size."""
        best_dist, best_rock = x_size * 2, -1
        for idx, rock in enumerate(rock_state.rocks):
            if rock.status != 0 and rock.count >= 0:
                d = Grid.manhattan_distance(rock_state.agent_pos, rock.pos)
                if d < best_dist:
                    best_dist, best_rock = d, idx
        return best_rock

    def _compute_prob(self, action, next_state, ob):
        """rock.py:250-264.  Scalar: floats.  Batched: float64 tensor (action/ob tensors)."""
        if self._scalar:                              # the same kernel, one particle: next_state in the reference's dict format
            return float(self.observation_prob([int(action)], self._pack_ref(next_state), [int(ob)])[0])
        return self.observation_prob(action, next_state, ob)


class StochasticRockEnv(RockEnv):
    """rock.py:428-504: every action only takes effect with probability ``p_move``; walls and
    empty samples cost nothing and never terminate."""
    _stochastic = True

    def __init__(self, board_size=7, num_rocks=8, use_heuristic=False, p_move=.8, batch_size=None, device="cuda",
                 seed=None, global_offset=0, track_belief_stats=False, track_history=False, history_next_is_reward=False):
        super().__init__(board_size, num_rocks, use_heuristic, batch_size, device, seed, global_offset, p_move,
                         track_belief_stats, track_history, history_next_is_reward)
        self.p_move = p_move
