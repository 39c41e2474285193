"""Tag on the GPU: host side of ``pomdp_tag_step`` / ``pomdp_tag_reset``.

Stands in for gym_pomdp/envs/tag.py ``TagEnv`` (84-280).  Packed state: bits 0-4 agent
cell, 5 bits per opponent cell from bit 5, bits 25-30 ``num_opp`` (signed), bit 31 done.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..geometry import TagGrid
from ..spaces import Discrete
from .base import BatchedPomdpEnv

TAG = 4  # tag.py:28-33


class TagState(object):
    """tag.py:284-291"""

    def __init__(self, coord):
        self.agent_pos = coord
        self.opponent_pos = []
        self.num_opp = 0

    def __str__(self):
        return str(len(self.opponent_pos))


class TagEnv(BatchedPomdpEnv):
    kind = _lib.KIND_TAG
    _abi = "tag"

    def __init__(self, num_opponents=1, move_prob=.8, obs_cells=29, board_size=(10, 5), batch_size=None,
                 device="cuda", seed=None, global_offset=0):
        super().__init__(batch_size, device, seed, global_offset)
        if obs_cells != 29 or tuple(board_size) != (10, 5):
            # TagGrid hard-codes the 29-cell board whatever these say (tag.py:46-66)
            raise ValueError("the Tag board is the reference's fixed 29-cell one")
        self.num_opponents = num_opponents
        self.move_prob = move_prob
        self._params = _lib.TagParams(num_opponents, 0, float(move_prob))
        self._reward_range = 10 * num_opponents
        self._discount = .95
        self.action_space = Discrete(5)
        self.grid = TagGrid(board_size, obs_cells=obs_cells)
        self.observation_space = Discrete(self.grid.n_tiles + 1)
        self.time = 0
        if not 1 <= num_opponents <= 4:
            raise ValueError("num_opponents must be in 1..4")
        L = _lib.lib()
        host = np.zeros(L.pomdp_tag_table_bytes(), dtype=np.uint8)
        _lib.check(L.pomdp_tag_build_table(host.ctypes.data), "pomdp_tag_build_table")
        self._table = torch.from_numpy(host).to(self.device)     # board maps, staged to shared memory by TMA per CTA

    def _c_head(self):
        return (ctypes.byref(self._params), _lib.ptr(self._table))

    def _c_step(self, state, action, next_state, obs, reward, flags, n, ctr):
        _lib.check(_lib.lib().pomdp_tag_step(
            ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state), _lib.ptr(obs),
            _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, self._stream()), "pomdp_tag_step")

    def _c_step_hist(self, state, action, next_state, obs, reward, flags, n, ctr, sink):
        _lib.check(_lib.lib().pomdp_tag_step_hist(
            ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state), _lib.ptr(obs),
            _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, ctypes.byref(sink), self._stream()),
            "pomdp_tag_step_hist")

    def _c_reset(self, state, obs, mask, n, ctr):
        _lib.check(_lib.lib().pomdp_tag_reset(
            ctypes.byref(self._params), _lib.ptr(state), _lib.ptr(obs), _lib.ptr(mask), n, self.global_offset,
            self._seed, ctr, self._stream()), "pomdp_tag_reset")

    def _hist_args(self):
        return 0, 0

    # ---------------------------------------------------------------------- codec ---
    def pack(self, agent, opp, num_opp=None, done=None):
        """agent int[n] cell ids, opp int[n, num_opponents] -> packed int32[n]."""
        agent = torch.as_tensor(agent, device=self.device).to(torch.int64)
        opp = torch.as_tensor(opp, device=self.device).to(torch.int64).reshape(agent.shape[0], self.num_opponents)
        v = agent.clone()
        for j in range(self.num_opponents):
            v |= opp[:, j] << (5 + 5 * j)
        nop = torch.full_like(agent, self.num_opponents) if num_opp is None else \
            torch.as_tensor(num_opp, device=self.device).to(torch.int64)
        v |= (nop & 63) << 25
        if done is not None:
            v |= torch.as_tensor(done, device=self.device).to(torch.int64) << 31
        return ((v + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)

    def unpack(self, words):
        """packed -> (agent[n], opp[n, num_opponents], num_opp[n], done[n])"""
        v = words.to(torch.int64) & 0xFFFFFFFF
        agent = v & 31
        opp = torch.stack([(v >> (5 + 5 * j)) & 31 for j in range(self.num_opponents)], dim=1)
        nop = (v >> 25) & 63
        nop = torch.where(nop >= 32, nop - 64, nop)
        return agent.to(torch.int32), opp.to(torch.int32), nop.to(torch.int32), ((v >> 31) & 1).bool()

    def encode_array(self, words):
        """int32 ``[agent_idx, opp_idx...]`` rows, the reference's ``_encode_state`` (tag.py:158-165)."""
        agent, opp, _, _ = self.unpack(words)
        return torch.cat([agent[:, None], opp], dim=1)

    def decode_array(self, arr):
        """inverse of ``encode_array`` (tag.py:167-179: every opponent with idx > -1 counts)."""
        arr = torch.as_tensor(arr, device=self.device).to(torch.int64).reshape(-1, 1 + self.num_opponents)
        return self.pack(arr[:, 0], arr[:, 1:], num_opp=(arr[:, 1:] > -1).sum(dim=1))

    # ---------------------------------------------------------------- scalar mode ---
    def _on_reset(self):
        self.time = 0
        self.last_action = 4

    def _state_to_ref(self, words):
        v = words[0]
        st = TagState(self.grid.get_tag_coord(v & 31))
        st.opponent_pos = [self.grid.get_tag_coord((v >> (5 + 5 * j)) & 31) for j in range(self.num_opponents)]
        nop = (v >> 25) & 63
        st.num_opp = nop - 64 if nop >= 32 else nop
        return st

    def _state_from_ref(self, state):
        agent = [self.grid.get_index(state.agent_pos)]
        opp = [[self.grid.get_index(o) for o in state.opponent_pos]]
        return self.pack(agent, opp, num_opp=[state.num_opp])

    def _after_scalar_step(self, action, ob):
        self.time += 1

    def _generate_legal(self, state=None):
        """tag.py:228-229"""
        if self._scalar and state is None:
            return list(range(self.action_space.n))
        return self.legal_mask(state)

    # ------------------------------------------------------------ heuristic action sets ---
    def _preferred_call(self, fn_name, state, last_obs, last_action, out, *tail):
        n = state.shape[0]
        conv = lambda t: None if t is None else torch.as_tensor(t, device=state.device).to(torch.int32).expand(n).contiguous()
        last_obs, last_action = conv(last_obs), conv(last_action)
        with self._guard():
            _lib.check(getattr(_lib.lib(), fn_name)(ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state.contiguous()),
                                                   _lib.ptr(last_obs), _lib.ptr(last_action), _lib.ptr(out), n, *tail,
                                                   self._stream()), fn_name)
        return out

    def preferred_mask_words(self, state=None, last_obs=None, last_action=None):
        """``_generate_preferred(history)`` (tag.py:231-243) for every particle as a bit mask over the five actions,
        int32[n].  The set depends on the history's last (ob, action) only; ``last_action`` None (or < 0) = empty history."""
        state = self.state if state is None else state
        return self._preferred_call("pomdp_tag_preferred_mask", state, last_obs, last_action,
                                    self._empty((state.shape[0],), torch.int32))

    def sample_preferred_actions(self, state=None, last_obs=None, last_action=None, out=None, step_ctr=None):
        """Batched ``np.random.choice(env._generate_preferred(history))``: int32[n], the POLICY draw of ``step_ctr``."""
        state = self.state if state is None else state
        action = self._empty((state.shape[0],), torch.int32) if out is None else out
        ctr = ((self._step_ctr + 1) & 0xFFFFFFFF) if step_ctr is None else int(step_ctr)
        return self._preferred_call("pomdp_tag_policy_preferred", state, last_obs, last_action, action, self.global_offset,
                                    self._seed, ctr)

    def _has_preferred_kernel(self):
        return True

    def _c_rollout_preferred(self, state, final_state, ret, steps, flags, n, ctr, max_steps, discount, first_action=None,
                             last_obs=None, last_action=None):
        """tag.py:303-316 with ``_generate_preferred``, fused.  ``last_obs`` / ``last_action`` int32[n] tensors: the
        history's last entry at the start (updated in place); None = an empty history."""
        _lib.check(_lib.lib().pomdp_tag_rollout_preferred(
            ctypes.byref(self._params), _lib.ptr(self._table), _lib.ptr(state), _lib.ptr(last_obs), _lib.ptr(last_action),
            _lib.ptr(first_action), _lib.ptr(final_state), _lib.ptr(ret), _lib.ptr(steps), _lib.ptr(flags), n,
            self.global_offset, self._seed, ctr, int(max_steps), float(discount), self._stream()), "pomdp_tag_rollout_preferred")

    def _generate_preferred(self, history):
        """tag.py:231-243.  Scalar mode: ``history`` is the caller's History (``.size``, ``[-1].ob``, ``[-1].action``);
        returns the reference's list.  Batched: ``history`` = (last_obs, last_action) int32[n] tensors or None for an empty
        history; returns bool[n, 5]."""
        if not self._scalar:
            lo, la = (None, None) if history is None else history
            words = self.preferred_mask_words(None, lo, la).to(torch.int64)
            a = torch.arange(self.action_space.n, device=words.device)
            return ((words[:, None] >> a) & 1).bool()
        if history.size == 0:
            return self._generate_legal()
        state = self._io_state.to(self.device) if self._io_state.device != self.device else self._io_state
        word = int(self.preferred_mask_words(state.reshape(-1), [int(history[-1].ob)], [int(history[-1].action)])[0])
        actions = [a for a in range(self.action_space.n) if (word >> a) & 1]
        assert len(actions) > 0
        return actions

    def _compute_prob(self, action, next_state, ob):
        """tag.py:209-217"""
        if self._scalar:                              # the same kernel, one particle: next_state is a TagState
            return float(self.observation_prob([int(action)], self._state_from_ref(next_state), [int(ob)])[0])
        return self.observation_prob(action, next_state, ob)
