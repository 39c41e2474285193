"""Tag's whole-transition table (include/pomdp_b200.h: pomdp_tag_build_table; layout in pomdp_core.h: TagTables.lut): every
entry decoded in numpy and compared with the oracle's restatement of tag.py:108-143 / 201-207 / 219-226 for every
(agent, opponent, action), both outcomes of the move coin and all four values of the pick bits.  The table is built on
the host, so the product library is checked here without a GPU, next to the test vehicle's."""
import ctypes

import numpy as np
import pytest

from gym_pomdp_b200 import build
from backends import build_hostsim
from oracle import pomdp_oracle as O

PITCH, CELLS = 33, 29
PLANE = 32 * PITCH
LUT_OFF = PLANE + 32            # words: pair[32 * 33], mv[32], then lut[5 * 32 * 33]


@pytest.mark.parametrize("which", ["product", "hostsim"])
def test_every_entry_is_the_reference_transition(which):
    L = ctypes.CDLL(build.OUT if which == "product" else build_hostsim())
    L.pomdp_tag_table_bytes.restype = ctypes.c_int64
    nbytes = L.pomdp_tag_table_bytes()
    assert nbytes == 4 * (LUT_OFF + 5 * PLANE) and nbytes % 16 == 0
    tab = np.zeros(nbytes // 4, np.uint32)
    assert L.pomdp_tag_build_table(ctypes.c_void_p(tab.ctypes.data)) == 0
    lut = tab[LUT_OFF:].reshape(5, 32, PITCH)
    valid = np.zeros(lut.shape, bool)
    T = O.bern_threshold(0.8)
    for a in range(5):
        for opp in range(CELLS):
            ox, oy = O.tag_get_coord(opp)
            for agent in range(CELLS):
                ax, ay = O.tag_get_coord(agent)
                e = int(lut[a, opp, agent])
                valid[a, opp, agent] = True
                assert e != 0
                for top2 in range(4):
                    for coin, w in ((True, top2 << 14), (False, (T & 0xFFFF0000) + 0x10000 + (top2 << 14))):
                        assert (w < T) == coin and (w >> 14) & 3 == top2
                        ax2, ay2, opps2, nop2, ob, rw, done = O.tag_step(ax, ay, [(ox, oy)], 1, a, lambda j, w=w: w)
                        d_opp = (e >> (5 * top2)) & 31 if coin else 0
                        assert opp ^ d_opp == O.tag_get_index(*opps2[0]), (a, opp, agent, top2, coin)
                        assert agent ^ ((e >> 20) & 31) == O.tag_get_index(ax2, ay2)
                        assert ((~e >> 25) & 31) == ob
                        is_tag, hit = (e >> 30) & 1, e >> 31
                        assert is_tag == (a == 4) and hit == (a == 4 and opp == agent)
                        assert (10. if hit else -10. if is_tag else -1.) == rw
                        assert (nop2 == 0) == bool(hit) == bool(done)
    assert not lut[~valid].any()                      # cell ids off the board (and the pitch column) read 0: flagged BAD_STATE


def test_one_word_serves_the_move_coin_and_the_move_choice():
    """The draw contract of Tag's opponent move (pomdp_core.h: tag_pick_word): the coin reads the word whole (w < T), the
    choice reads bits 15-14 (= floor(u' * len) for the word w << 16 and len in {2, 4}).  Counted exactly over all 2^32
    words: given that the opponent moves, every pick has probability 1/len to within 2^14 / T -- the two outcomes are
    independent far below anything 10^7 draws can resolve -- and the marginal of the coin is T / 2^32 exactly."""
    for move_prob in (0.8, 0.5, 0.05, 1.0, 1e-4):
        T = O.bern_threshold(move_prob)
        full, rest = divmod(T, 1 << 16)
        counts = [full * (1 << 14) + min(max(rest - v * (1 << 14), 0), 1 << 14) for v in range(4)]      # w < T with bits 15-14 == v
        assert sum(counts) == T
        for v in range(4):
            assert abs(counts[v] / T - 0.25) <= (1 << 14) / T
        for first in (0, 1):                                                                              # len == 2: bit 15 alone
            assert abs((counts[2 * first] + counts[2 * first + 1]) / T - 0.5) <= (1 << 15) / T
    # the rule the oracle states it with
    for w in (0, 1 << 14, 3 << 14, 0xFFFFFFFF, 0x12345678):
        assert O.rand_below(O.tag_pick_word(w), 4) == (w >> 14) & 3 and O.rand_below(O.tag_pick_word(w), 2) == (w >> 15) & 1
