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

    def __init__(self, n, k, device):
        self.count = torch.zeros((n, k), dtype=torch.int32, device=device)
        self.measured = torch.zeros((n, k), dtype=torch.int32, device=device)
        self.lkv = torch.ones((n, k), dtype=torch.float64, device=device)
        self.lkw = torch.ones((n, k), dtype=torch.float64, device=device)
        self.prob_valuable = torch.full((n, k), .5, dtype=torch.float64, device=device)

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


class RockEnv(BatchedPomdpEnv):
    kind = _lib.KIND_ROCK
    _abi = "rock"
    _stochastic = False

    def __init__(self, board_size=7, num_rocks=8, use_heuristic=False, batch_size=None, device="cuda", seed=0,
                 global_offset=0, p_move=0.8, track_belief_stats=False):
        super().__init__(batch_size, device, seed, global_offset)
        self.track_belief_stats = bool(track_belief_stats)   # batched mode: keep rock.py's per-rock side-stats current
        self.belief_stats = None
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
        self.action_space = Discrete(5 + num_rocks)                   # rock.py:113
        self.observation_space = Discrete(3)                          # rock.py:114
        self._discount = .95
        self._reward_range = 20
        self._penalization = 0 if self._stochastic else -100
        self._query = 0
        self._side = None   # scalar mode: per-rock belief side-stats (rock.py:82-86)

    # -------------------------------------------------------------------- C calls ---
    def _c_head(self):
        return (ctypes.byref(self._params), _lib.ptr(self._table))

    _c_query_head = _c_head

    def _c_step(self, state, action, next_state, obs, reward, flags, n, ctr):
        _lib.check(_lib.lib().pomdp_rock_step(
            ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state),
            _lib.ptr(obs), _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, self._stream()),
            "pomdp_rock_step")

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
        n = action.shape[0]
        action = torch.as_tensor(action, device=self.device).to(torch.int32).contiguous()
        obs = torch.as_tensor(obs, device=self.device).to(torch.int32).contiguous()
        with self._guard():
            _lib.check(_lib.lib().pomdp_rock_belief_update(
                ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(next_state), _lib.ptr(action), _lib.ptr(obs),
                _lib.ptr(stats.count), _lib.ptr(stats.measured), _lib.ptr(stats.lkv), _lib.ptr(stats.lkw),
                _lib.ptr(stats.prob_valuable), n, self._stream()), "pomdp_rock_belief_update")
        return stats

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
    def reset(self, mask=None):
        obs = super().reset(mask)
        if self.track_belief_stats and not self._scalar:
            if self.belief_stats is None:
                self.belief_stats = self.new_belief_stats()
            else:
                self.belief_stats.reset(mask)
        return obs

    def step(self, action):
        out = super().step(action)
        if self.track_belief_stats and not self._scalar:
            self.update_belief_stats(self.belief_stats, self.state, action, out[0])
            out[3]["belief_stats"] = self.belief_stats
        return out

    def _on_reset(self):
        self._query = 0
        self.last_action = SAMPLE
        if self._scalar:
            self._side = [dict(count=0, measured=0, lkw=1., lkv=1., prob_valuable=.5) for _ in range(self.num_rocks)]

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

    def _state_from_ref(self, state):
        ax, ay = state["agent_pos"]
        status = [[int(r["status"]) for r in state["rocks"]]]
        self._side = [dict(count=r.get("count", 0), measured=r.get("measured", 0), lkw=r.get("lkw", 1.),
                           lkv=r.get("lkv", 1.), prob_valuable=r.get("prob_valuable", .5)) for r in state["rocks"]]
        return self.pack([ax], [ay], status)

    def _reward_to_py(self, reward, action):
        return int(reward)

    def _raise_for_flags(self, flags):
        if flags & _lib.FLAG_BAD_STATE:
            # the reference dies with IndexError at rock.py:162 on a dangling grid id
            raise IndexError("list index out of range")

    @staticmethod
    def _efficiency(agent_pos, rock_pos, hed=20):
        d = Grid.euclidean_distance(agent_pos, rock_pos)
        return (1 + pow(2, -d / hed)) * .5

    def _after_scalar_step(self, action, ob):
        self._query += 1
        if action > SAMPLE and ob != NULL:        # rock.py:177-191: belief side-stats of the checked rock
            r = self._side[action - SAMPLE - 1]
            x, y, _ = self._decode_py(self._host_words())
            eff = self._efficiency((x, y), self._rock_pos[action - SAMPLE - 1])
            r["measured"] += 1
            if ob == GOOD:
                r["count"] += 1
                r["lkv"] *= eff
                r["lkw"] *= (1 - eff)
            else:
                r["count"] -= 1
                r["lkw"] *= eff
                r["lkv"] *= (1 - eff)
            denom = (.5 * r["lkv"]) + (.5 * r["lkw"])
            r["prob_valuable"] = (.5 * r["lkv"]) / denom if denom else float("nan")

    # ------------------------------------------------------------- planner hooks ---
    def _generate_legal(self, state=None):
        """rock.py:273-291.  Scalar mode: the reference's list (same order).  Batched: bool
        mask [n, n_actions] (a set; the reference's duplicate entries for Rock(15,15)'s
        doubled rock collapse)."""
        if self._scalar and state is None:
            x, y, status = self._decode_py(self._host_words())
            n = self.grid.x_size
            legal = [1]
            if y + 1 < n:
                legal.append(0)
            if y - 1 >= 0:
                legal.append(2)
            if x - 1 >= 0:
                legal.append(3)
            rock = int(self._grid_map[x | (y << 4)])
            if rock >= 0 and int(status[rock]) != 0:
                legal.append(SAMPLE)
            for i in range(self.num_rocks):
                if int(status[i]) != 0:
                    p = self._rock_pos[i]
                    legal.append(int(self._grid_map[p.x | (p.y << 4)]) + 1 + SAMPLE)
            return legal
        return self.legal_mask(state)

    def _generate_preferred(self, history):
        if not self._use_heuristic:
            return self._generate_legal()
        raise NotImplementedError("use_heuristic=True rollouts are not on the device path yet (SURVEY.md §8f rank 3)")

    def _compute_prob(self, action, next_state, ob):
        """rock.py:250-264.  Scalar: floats.  Batched: float64 tensor (action/ob tensors)."""
        if self._scalar:
            if action <= SAMPLE:
                return int(ob == NULL)
            rock = next_state["rocks"][action - SAMPLE - 1]
            eff = self._efficiency(next_state["agent_pos"], rock["pos"])
            if (ob == GOOD and rock["status"] == 1) or (ob == BAD and rock["status"] == -1):
                return eff
            return 1 - eff
        return self.observation_prob(action, next_state, ob)


class StochasticRockEnv(RockEnv):
    """rock.py:428-504: every action only takes effect with probability ``p_move``; walls and
    empty samples cost nothing and never terminate."""
    _stochastic = True

    def __init__(self, board_size=7, num_rocks=8, use_heuristic=False, p_move=.8, batch_size=None, device="cuda",
                 seed=0, global_offset=0, track_belief_stats=False):
        super().__init__(board_size, num_rocks, use_heuristic, batch_size, device, seed, global_offset, p_move,
                         track_belief_stats)
        self.p_move = p_move
