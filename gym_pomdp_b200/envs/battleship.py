"""BattleShip on the GPU: host side of ``pomdp_battleship_step`` / ``pomdp_battleship_reset``.

Stands in for gym_pomdp/envs/battleship.py ``BattleShipEnv`` (64-211).  Packed state: 8
int32 words per board -- words 0-3 occupied bits (cell c = x_size*y + x, i.e. the action
index), word 3 bits 24-30 ``total_remaining`` and bit 31 done, words 4-7 visited bits.
Unlike the reference (battleship.py:11 "TODO fix state", 124-126) the board travels WITH
the state, so ``_set_state`` restores a consistent game.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..geometry import Grid
from ..spaces import Discrete
from .base import BatchedPomdpEnv


class ShipState(object):
    """battleship.py:40-43, plus the board the reference keeps in ``env.grid``."""

    def __init__(self):
        self.ships = []
        self.total_remaining = 0
        self.occupied = None   # bool [x_size, y_size]
        self.visited = None


class BattleShipEnv(BatchedPomdpEnv):
    kind = _lib.KIND_BATTLESHIP
    _abi = "battleship"
    state_words = 8

    def __init__(self, board_size=(5, 5), max_len=3, batch_size=None, device="cuda", seed=None, global_offset=0,
                 reset_mode="table"):
        super().__init__(batch_size, device, seed, global_offset)
        self.grid = Grid(*board_size)
        self._params = _lib.BattleshipParams(board_size[0], board_size[1], max_len, 0)
        self.action_space = Discrete(self.grid.n_tiles)
        self.observation_space = Discrete(2)
        self.num_obs = 2
        self._reward_range = self.action_space.n / 4.
        self._discount = 1.
        self.total_remaining = max_len - 1      # battleship.py:74 (an env attribute the reference only asserts on)
        self.max_len = max_len + 1              # battleship.py:75
        if reset_mode not in ("table", "scan", "warpscan", "rejection"):
            raise ValueError("reset_mode must be 'table' (placement tables, thread per env), 'scan' (bitboard scan, thread per "
                             "env), 'warpscan' (warp per env) -- all three give the same boards -- or 'rejection' (the "
                             "reference's loop)")
        self.reset_mode = reset_mode
        self.reset_flags = None
        self.t = 0
        self.tot_rw = 0
        L = _lib.lib()
        nbytes = L.pomdp_battleship_table_bytes(ctypes.byref(self._params))
        if nbytes < 0:
            raise ValueError(L.pomdp_last_error().decode())
        host = np.zeros(nbytes, dtype=np.uint8)
        _lib.check(L.pomdp_battleship_build_table(ctypes.byref(self._params), host.ctypes.data), "pomdp_battleship_build_table")
        self._table = torch.from_numpy(host).to(self.device)      # accepted placements of ships 0 and 1 (static per config)

    def _c_step(self, state, action, next_state, obs, reward, flags, n, ctr):
        _lib.check(_lib.lib().pomdp_battleship_step(
            ctypes.byref(self._params), _lib.ptr(state), _lib.ptr(action), _lib.ptr(next_state), _lib.ptr(obs),
            _lib.ptr(reward), _lib.ptr(flags), n, self._stream()), "pomdp_battleship_step")

    def _c_reset(self, state, obs, mask, n, ctr):
        L = _lib.lib()
        # every reset env gets its flag word written by the kernel; only a masked reset needs the others pre-cleared
        self.reset_flags = (torch.empty if mask is None else torch.zeros)(n, dtype=torch.int32, device=self.device)
        tail = (_lib.ptr(state), _lib.ptr(obs), _lib.ptr(self.reset_flags), _lib.ptr(mask), n, self.global_offset,
                self._seed, ctr, self._stream())
        if self.reset_mode in ("table", "scan"):
            table = _lib.ptr(self._table) if self.reset_mode == "table" else None
            rc = L.pomdp_battleship_reset(ctypes.byref(self._params), table, *tail)
        else:
            fn = L.pomdp_battleship_reset_warpscan if self.reset_mode == "warpscan" else L.pomdp_battleship_reset_rejection
            rc = fn(ctypes.byref(self._params), *tail)
        _lib.check(rc, "pomdp_battleship_reset")

    def reset(self, mask=None, seed=None, options=None):
        """battleship.py:131-137.  A (board, max_len) with no legal placement makes the reference's rejection loop spin
        forever; here the kernels flag it: scalar mode raises, batched mode ORs FLAG_BAD_STATE into ``flags``."""
        obs = super().reset(mask, seed, options)
        rf = self.reset_flags
        if rf is not None:
            if self._scalar:
                if int(rf[0]) & _lib.FLAG_BAD_STATE:
                    raise RuntimeError("BattleShip: no legal ship placement exists on this board (the reference would loop forever)")
            else:
                self.flags = self.flags | rf
        return obs

    def _hist_args(self):
        return self.grid.n_tiles, 0

    # ---------------------------------------------------------------------- codec ---
    def pack(self, occupied, visited, total_remaining=None, done=None):
        """occupied / visited bool [n, x_size, y_size] -> packed int32 [n, 8]."""
        occ = torch.as_tensor(occupied, device=self.device).bool()
        vis = torch.as_tensor(visited, device=self.device).bool()
        n = occ.shape[0]
        # cell c = x_size * y + x  ->  flatten in (y, x) order
        occ_c = occ.permute(0, 2, 1).reshape(n, -1).to(torch.int64)
        vis_c = vis.permute(0, 2, 1).reshape(n, -1).to(torch.int64)
        words = torch.zeros((n, 8), dtype=torch.int64, device=self.device)
        for c in range(self.grid.n_tiles):
            words[:, c >> 5] |= occ_c[:, c] << (c & 31)
            words[:, 4 + (c >> 5)] |= vis_c[:, c] << (c & 31)
        rem = (occ_c * (1 - vis_c)).sum(dim=1) if total_remaining is None else \
            torch.as_tensor(total_remaining, device=self.device).to(torch.int64)
        words[:, 3] |= (rem & 0x7F) << 24
        if done is not None:
            words[:, 3] |= torch.as_tensor(done, device=self.device).to(torch.int64) << 31
        return ((words + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)

    def unpack(self, words):
        """packed -> (occupied[n, X, Y], visited[n, X, Y], total_remaining[n], done[n])"""
        w = words.to(torch.int64) & 0xFFFFFFFF
        n = w.shape[0]
        c = torch.arange(self.grid.n_tiles, device=words.device)
        occ_c = (w[:, (c >> 5)] >> (c & 31)) & 1
        vis_c = (w[:, 4 + (c >> 5)] >> (c & 31)) & 1
        X, Y = self.grid.x_size, self.grid.y_size
        occ = occ_c.reshape(n, Y, X).permute(0, 2, 1).bool()
        vis = vis_c.reshape(n, Y, X).permute(0, 2, 1).bool()
        return occ, vis, ((w[:, 3] >> 24) & 0x7F).to(torch.int32), ((w[:, 3] >> 31) & 1).bool()

    # ---------------------------------------------------------------- scalar mode ---
    def _on_reset(self):
        self.tot_rw = 0
        self.t = 0
        self.last_action = -1

    def _state_to_ref(self, words):
        X, Y = self.grid.x_size, self.grid.y_size
        occ_bits = words[0] | (words[1] << 32) | (words[2] << 64) | ((words[3] & 0x00FFFFFF) << 96)
        vis_bits = words[4] | (words[5] << 32) | (words[6] << 64) | (words[7] << 96)
        st = ShipState()
        st.total_remaining = (words[3] >> 24) & 0x7F
        st.occupied = np.array([[(occ_bits >> (X * y + x)) & 1 for y in range(Y)] for x in range(X)], dtype=bool)
        st.visited = np.array([[(vis_bits >> (X * y + x)) & 1 for y in range(Y)] for x in range(X)], dtype=bool)
        return st

    def _state_from_ref(self, state):
        return self.pack(np.asarray(state.occupied)[None], np.asarray(state.visited)[None],
                         total_remaining=[state.total_remaining])

    def _reward_to_py(self, reward, action):
        return int(reward)

    def _after_scalar_step(self, action, ob):
        self.t += 1

    def _step_scalar(self, action):
        out = super()._step_scalar(action)
        self.tot_rw += out[1]
        return out

    def _generate_legal(self, state=None):
        """battleship.py:157-165: the unvisited cells."""
        if self._scalar and state is None:
            w = self._host_words()
            vis_bits = w[4] | (w[5] << 32) | (w[6] << 64) | (w[7] << 96)
            return [a for a in range(self.grid.n_tiles) if not (vis_bits >> a) & 1]     # action index = x_size * y + x
        return self.legal_mask(None if state is None else state.reshape(-1, 8))

    def _generate_preferred(self, history):
        return self._generate_legal()

    def _compute_prob(self, action, next_state, ob):
        """battleship.py:80-89"""
        if self._scalar:                              # the same kernel, one particle: next_state is a ShipState
            return float(self.observation_prob([int(action)], self._state_from_ref(next_state), [int(ob)])[0])
        return self.observation_prob(action, next_state, ob)
