"""Tiger on the GPU: host side of ``pomdp_tiger_step`` / ``pomdp_tiger_reset``.

Stands in for gym_pomdp/envs/tiger.py ``TigerEnv`` (47-172).  Packed state: bit 0 = the
door the tiger is behind (0 left, 1 right), bit 31 done.
"""
import ctypes

import torch

from .. import _lib
from ..spaces import Discrete
from .base import BatchedPomdpEnv

LEFT, RIGHT, LISTEN = 0, 1, 2   # tiger.py:21-24; observations: 0 left, 1 right, 2 null (tiger.py:10-13)


class TigerEnv(BatchedPomdpEnv):
    kind = _lib.KIND_TIGER
    _abi = "tiger"

    def __init__(self, seed=None, correct_prob=.85, batch_size=None, device="cuda", global_offset=0):
        super().__init__(batch_size, device, seed, global_offset)
        self.correct_prob = correct_prob
        # the reference's _sample_ob never reads self.correct_prob: its default argument .85 is
        # what is used (tiger.py:141, called at tiger.py:86)
        self._params = _lib.TigerParams(0.85)
        self.action_space = Discrete(3)
        self.state_space = Discrete(2)
        self.observation_space = Discrete(3)
        self._discount = .95
        self._reward_range = 10
        self._query = 0
        self.t = 0

    def _c_step(self, state, action, next_state, obs, reward, flags, n, ctr):
        _lib.check(_lib.lib().pomdp_tiger_step(
            ctypes.byref(self._params), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state), _lib.ptr(obs),
            _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, self._stream()),
            "pomdp_tiger_step")

    def _c_step_hist(self, state, action, next_state, obs, reward, flags, n, ctr, sink):
        _lib.check(_lib.lib().pomdp_tiger_step_hist(
            ctypes.byref(self._params), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state), _lib.ptr(obs),
            _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, ctypes.byref(sink), self._stream()),
            "pomdp_tiger_step_hist")

    def _c_reset(self, state, obs, mask, n, ctr):
        _lib.check(_lib.lib().pomdp_tiger_reset(
            ctypes.byref(self._params), _lib.ptr(state), _lib.ptr(obs), _lib.ptr(mask), n, self.global_offset,
            self._seed, ctr, self._stream()), "pomdp_tiger_reset")

    def _hist_args(self):
        return 0, 0

    def pack(self, tiger, done=None):
        v = torch.as_tensor(tiger, device=self.device).to(torch.int64) & 1
        if done is not None:
            v = v | (torch.as_tensor(done, device=self.device).to(torch.int64) << 31)
        return ((v + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)

    def unpack(self, words):
        v = words.to(torch.int64) & 0xFFFFFFFF
        return (v & 1).to(torch.int32), ((v >> 31) & 1).bool()

    def _on_reset(self):
        self.t = 0
        self._query = 0
        self.last_action = LISTEN

    def _state_to_ref(self, words):
        return words[0] & 1          # tiger.py:107-109: the state is a plain int

    def _state_from_ref(self, state):
        return self.pack([int(state)])

    def _reward_to_py(self, reward, action):
        return int(reward)

    def _after_scalar_step(self, action, ob):
        self.t += 1
        self._query += 1

    def _generate_legal(self, state=None):
        if self._scalar and state is None:
            return list(range(self.action_space.n))
        return self.legal_mask(state)

    def _generate_preferred(self, history):
        return self._generate_legal()

    def _compute_prob(self, action, next_state, ob, correct_prob=.85):
        """tiger.py:125-138"""
        if self._scalar:                              # the same kernel, one particle
            return float(self.observation_prob([int(action)], self._state_from_ref(next_state), [int(ob)], float(correct_prob))[0])
        return self.observation_prob(action, next_state, ob, float(correct_prob))
