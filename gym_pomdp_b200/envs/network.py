"""Network on the GPU: host side of ``pomdp_network_step`` / ``pomdp_network_reset``.

Stands in for gym_pomdp/envs/network.py ``NetworkEnv`` (24-168).  Packed state: bit m =
machine m is up; bit 31 done (never set -- the reference's episode never ends).
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..spaces import Discrete
from .base import BatchedPomdpEnv

OFF, ON, NULL = 0, 1, 2   # network.py:12-15


class NetworkEnv(BatchedPomdpEnv):
    kind = _lib.KIND_NETWORK
    _abi = "network"
    _reward_unit = 10          # packed results carry the reward in tenths (network.py:104, 108)

    def __init__(self, n_machines=10, problem_type=3, depth=60, batch_size=None, device="cuda", seed=None,
                 global_offset=0):
        super().__init__(batch_size, device, seed, global_offset)
        self._p = 0.1
        self._q = 0.33
        self._query = 0
        self._depth = 60          # network.py:31 ignores the argument
        self._n_machines = n_machines
        self._params = _lib.NetworkParams(n_machines, problem_type, self._p, self._q, self._p_ob)
        self.action_space = Discrete(n_machines * 2 + 1)
        self.observation_space = Discrete(3)
        self._discount = .95
        self._reward_range = n_machines * 2
        self.neighbours = self.make_3legs_neighbours(n_machines) if problem_type == 3 else \
            self.make_ring_neighbours(n_machines)
        if n_machines > 30:
            raise ValueError("n_machines must fit one state word (<= 30)")
        self._server = 0
        self.t = 0

    @property
    def _p_ob(self):
        return .95

    def _c_step(self, state, action, next_state, obs, reward, flags, n, ctr):
        _lib.check(_lib.lib().pomdp_network_step(
            ctypes.byref(self._params), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state), _lib.ptr(obs),
            _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, self._stream()),
            "pomdp_network_step")

    def _c_step_hist(self, state, action, next_state, obs, reward, flags, n, ctr, sink):
        _lib.check(_lib.lib().pomdp_network_step_hist(
            ctypes.byref(self._params), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state), _lib.ptr(obs),
            _lib.ptr(reward), _lib.ptr(flags), n, self.global_offset, self._seed, ctr, ctypes.byref(sink), self._stream()),
            "pomdp_network_step_hist")

    def _c_reset(self, state, obs, mask, n, ctr):
        _lib.check(_lib.lib().pomdp_network_reset(
            ctypes.byref(self._params), _lib.ptr(state), _lib.ptr(obs), _lib.ptr(mask), n, self._stream()),
            "pomdp_network_reset")

    def _hist_args(self):
        return self._n_machines, 0

    # ---------------------------------------------------------------------- codec ---
    def pack(self, machines):
        """0/1 array [n, n_machines] (the reference's int8 state, network.py:67) -> packed int32[n]."""
        m = torch.as_tensor(machines, device=self.device).to(torch.int64).reshape(-1, self._n_machines)
        w = (1 << torch.arange(self._n_machines, device=self.device, dtype=torch.int64))
        return ((m != 0).to(torch.int64) * w).sum(dim=1).to(torch.int32)

    def unpack(self, words):
        """packed -> int8 [n, n_machines]"""
        v = words.to(torch.int64)[:, None]
        return ((v >> torch.arange(self._n_machines, device=words.device, dtype=torch.int64)) & 1).to(torch.int8)

    # ---------------------------------------------------------------- scalar mode ---
    def _on_reset(self):
        self.t = 0
        self._query = 0
        self.last_action = self._n_machines * 2
        self._server = 0

    def _state_to_ref(self, words):
        return np.array([(words[0] >> m) & 1 for m in range(self._n_machines)], dtype=np.int8)

    def _state_from_ref(self, state):
        return self.pack(np.asarray(state).reshape(1, -1))

    def _reward_to_py(self, reward, action):
        # the reference returns an int when the action is the no-op (network.py:75, 101) and
        # a float otherwise; tenths are exact in the kernel, so rebuild the reference's double
        tenths = int(round(reward * 10))
        return tenths // 10 if action == 2 * self._n_machines else tenths / 10.0

    def _after_scalar_step(self, action, ob):
        self.t += 1
        self._query += 1

    def _generate_legal(self, state=None):
        if self._scalar and state is None:
            return list(range(self.action_space.n))
        return self.legal_mask(state)

    def _generate_preferred(self, history):
        return self._generate_legal()

    def sample_action(self):
        return int(np.random.choice(self._generate_legal()))

    def _compute_prob(self, action, next_state, ob):
        """network.py:43-55"""
        if self._scalar:                              # the same kernel, one particle: next_state is the int8[n] machine array
            return float(self.observation_prob([int(action)], self._state_from_ref(next_state), [int(ob)])[0])
        return self.observation_prob(action, next_state, ob)

    @staticmethod
    def make_ring_neighbours(n_machines):
        return [[(i + 1) % n_machines, (i + n_machines - 1) % n_machines] for i in range(n_machines)]

    @staticmethod
    def make_3legs_neighbours(n_machines):
        assert n_machines >= 4 and n_machines % 3 == 1
        nb = [[1, 2, 3]] + [[] for _ in range(n_machines - 1)]
        for i in range(1, n_machines):
            if i < n_machines - 3:
                nb[i].append(i + 3)
            nb[i].append(0 if i <= 4 else i - 3)
        return nb
