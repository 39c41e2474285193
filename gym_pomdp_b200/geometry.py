"""Geometry helpers: scalar host mirrors and batched device versions.

The reference's geometry (gym_pomdp/envs/coord.py:7-114, tag.py:36-78) is a set of small
pure functions wrapped in classes.  The hot path never touches the classes here -- the
kernels carry coordinates as packed nibbles -- they exist so single-instance callers keep
getting ``Coord`` objects in ``info["state"]`` and can keep asking ``env.grid`` for
indices.  The ``*_batch`` functions run the same integer arithmetic on the GPU through
``pomdp_coord_op`` and are what the bit-exact parity tests exercise.
"""
import math
from collections import namedtuple
from enum import Enum

import numpy as np
import torch

from . import _lib

# (dx, dy) of N, E, S, W, NULL -- coord.py:101-106
MOVE_DELTAS = ((0, 1), (1, 0), (0, -1), (-1, 0), (0, 0))


class Coord(namedtuple("Coord", "x y")):
    """2-D integer point with component-wise ``+`` (coord.py:7-19)."""
    __slots__ = ()

    def __add__(self, other):
        return Coord(self[0] + other[0], self[1] + other[1])

    def is_valid(self):
        return min(self) >= 0

    def __str__(self):
        return "%d,%d" % self


class Moves(Enum):
    NORTH = Coord(*MOVE_DELTAS[0])
    EAST = Coord(*MOVE_DELTAS[1])
    SOUTH = Coord(*MOVE_DELTAS[2])
    WEST = Coord(*MOVE_DELTAS[3])
    NULL = Coord(*MOVE_DELTAS[4])

    @staticmethod
    def get_coord(idx):
        return Coord(*MOVE_DELTAS[idx])

    @staticmethod
    def sample():
        """coord.py:112-114"""
        return np.random.randint(len(Moves))


def opposite(move):
    """coord.py:75-77"""
    return (move + 2) % 4


class Grid(object):
    """Rectangular board: cell index = x_size * y + x (coord.py:58-66)."""

    def __init__(self, x_size=10, y_size=5):
        self.x_size, self.y_size = x_size, y_size
        self.n_tiles = x_size * y_size
        self.build_board()

    # the host-side ``board`` of the reference (coord.py:36-52): an int8 (x_size, y_size) array of -1 that callers index
    # with a Coord -- ``grid[pos]`` (rock.py:161); out-of-range reads give None, negative indices wrap as numpy's do.
    # The kernels do not read it: RockSample's rock-id map travels in the TMA-staged table.
    def __iter__(self):
        return iter(self.board)

    def __setitem__(self, idx, value):
        try:
            self.board[idx] = value
        except IndexError:
            raise IndexError()

    def __getitem__(self, idx):
        try:
            return self.board[idx]
        except IndexError:
            return None

    def build_board(self, value=1):
        self.board = np.zeros(self.get_size, dtype=np.int8) - value

    def sample(self):
        """coord.py:68-69"""
        return self.get_coord(np.random.randint(self.n_tiles))

    @property
    def get_size(self):
        return self.x_size, self.y_size

    def get_index(self, coord):
        return self.x_size * coord[1] + coord[0]

    def get_coord(self, idx):
        if not 0 <= idx < self.n_tiles:
            raise AssertionError(idx)
        y, x = divmod(idx, self.x_size)
        return Coord(x, y)

    def is_inside(self, coord):
        return 0 <= coord[0] < self.x_size and 0 <= coord[1] < self.y_size

    opposite = staticmethod(opposite)

    @staticmethod
    def euclidean_distance(c1, c2):
        """The reference's name; its value is the 1-norm (np.linalg.norm(., 1), coord.py:79-81)."""
        return float(abs(c1[0] - c2[0]) + abs(c1[1] - c2[1]))

    @staticmethod
    def manhattan_distance(c1, c2):
        """... and this one is the 2-norm (coord.py:83-85)."""
        return math.hypot(c1[0] - c2[0], c1[1] - c2[1])

    @staticmethod
    def directional_distance(c1, c2, d):
        """coord.py:87-98, as written there (direction 2 mixes c2.y with c1.x)."""
        if d == 0:
            return c1[1] - c2[1]
        elif d == 1:
            return c1[0] - c2[0]
        elif d == 2:
            return c2[1] - c1[0]
        elif d == 3:
            return c2[0] - c1[0]
        raise NotImplementedError()


class TagGrid(Grid):
    """The fixed 29-cell Tag board (tag.py:36-78): rows y=0,1 are 10 wide, then a 3x3 block."""
    N_CELLS = 29

    def __init__(self, board_size=(10, 5), obs_cells=29):
        Grid.__init__(self, *board_size)
        self.n_tiles = obs_cells

    def sample(self):
        """tag.py:43-44"""
        return self.get_tag_coord(np.random.randint(0, 29))

    def is_inside(self, coord):
        x, y = coord
        return (5 <= x < 8 and y < 5) if y >= 2 else (0 <= x < 10 and y >= 0)

    def get_tag_coord(self, idx):
        if not 0 <= idx < self.n_tiles:
            raise AssertionError(idx)
        if idx < 20:
            return Coord(idx % 10, idx // 10)
        q, r = divmod(idx - 20, 3)
        return Coord(r + 5, q + 2)

    def get_index(self, coord):
        x, y = coord
        if not (0 <= x < 10 and 0 <= y < 5 and (y < 2 or 5 <= x < 8)):
            raise AssertionError(coord)
        return y * 10 + x if y < 2 else 20 + (y - 2) * 3 + (x - 5)

    def is_corner(self, coord):
        x, y = coord
        if not self.is_inside(coord):
            return False
        return x in (0, 9) if y < 2 else (y == 4 and x in (5, 7))

    @property
    def get_available_coord(self):
        return [self.get_tag_coord(i) for i in range(self.n_tiles)]


# --------------------------------------------------------------------- device batch ---
def _coord_op(op, a, b, pair_out, x_size=0, y_size=0):
    if a.device.type != "cuda" and not _lib.is_hostsim():
        raise RuntimeError("gym_pomdp_b200 geometry kernels run on CUDA tensors only")
    a = a.to(torch.int32).contiguous()
    n = a.shape[0]
    out = torch.empty((n, 2) if pair_out else (n,), dtype=torch.int32, device=a.device)
    if b is not None:
        b = b.to(device=a.device, dtype=torch.int32).contiguous()
    L = _lib.lib()

    def call():
        _lib.check(L.pomdp_coord_op(op, x_size, y_size, _lib.ptr(a), _lib.ptr(b), _lib.ptr(out), n,
                                    _lib.stream_handle(a.device)), "pomdp_coord_op")
    if a.device.type == "cuda":
        with torch.cuda.device(a.device):
            call()
    else:
        call()
    return out


def grid_get_index_batch(coords, x_size):
    """coords int32[n,2] -> int32[n]"""
    return _coord_op(_lib.COORD_GET_INDEX, coords, None, False, x_size)


def grid_get_coord_batch(idx, x_size):
    """idx int32[n] -> int32[n,2]"""
    return _coord_op(_lib.COORD_GET_COORD, idx, None, True, x_size)


def grid_is_inside_batch(coords, x_size, y_size):
    return _coord_op(_lib.COORD_IS_INSIDE, coords, None, False, x_size, y_size).bool()


def coord_add_move_batch(coords, moves):
    """Coord + Moves.get_coord(m), elementwise"""
    return _coord_op(_lib.COORD_ADD_MOVE, coords, moves, True)


def l1_distance_batch(a, b):
    """Grid.euclidean_distance (the 1-norm), elementwise"""
    return _coord_op(_lib.COORD_L1, a, b, False)


def tag_get_index_batch(coords):
    return _coord_op(_lib.COORD_TAG_GET_INDEX, coords, None, False)


def tag_get_coord_batch(idx):
    return _coord_op(_lib.COORD_TAG_GET_COORD, idx, None, True)


def tag_is_inside_batch(coords):
    return _coord_op(_lib.COORD_TAG_IS_INSIDE, coords, None, False).bool()
