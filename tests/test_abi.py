"""The C-ABI boundary: libpomdp_b200.so loads on a GPU-less box, exports every symbol that
include/pomdp_b200.h declares (and nothing the header does not know), and its host-side
argument validation follows the error convention of the header.  No compute entry point
is launched here -- those are the ``-m gpu`` tests.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from gym_pomdp_b200 import _lib, build

from backends import build_hostsim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pomdp_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pomdp_[a-z0-9_]+)\s*\(", src)))


def header_defines():
    src = open(HEADER).read()
    return {k: int(v, 0) for k, v in re.findall(r"#define\s+(POMDP_[A-Z0-9_]+)\s+\(?(-?\w+)\)?", src)}


@pytest.fixture(scope="module")
def product():
    return ctypes.CDLL(build.build(force=False))


def test_header_declares_what_the_binding_binds():
    assert header_functions() == sorted(_lib.EXPORTED_SYMBOLS)


def test_product_library_exports_every_declared_symbol(product):
    for name in header_functions():
        assert hasattr(product, name), name
    # ... and is the CUDA build, not the host simulation
    assert not hasattr(product, "pomdp_is_hostsim")
    dyn = subprocess.run(["nm", "-D", "--defined-only", build.OUT], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (pomdp_[a-z0-9_]+)$", dyn, flags=re.M)))
    assert exported == header_functions()


def test_product_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", build.OUT], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, out


def test_product_library_uses_tma_bulk_copies():
    """UBLKCP = cp.async.bulk (TMA) in SASS: the Rock table and the BattleShip board tiles."""
    sass = subprocess.run(["cuobjdump", "-sass", build.OUT], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass
    assert "SYNCS" in sass          # mbarrier


def test_hostsim_exports_the_same_abi():
    lib = ctypes.CDLL(build_hostsim())
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.pomdp_is_hostsim() == 1


def test_constants_match_header():
    d = header_defines()
    assert d["POMDP_ABI_VERSION"] == _lib.ABI_VERSION
    assert (d["POMDP_FLAG_DONE"], d["POMDP_FLAG_BAD_ACTION"], d["POMDP_FLAG_STEPPED_DONE"], d["POMDP_FLAG_BAD_STATE"]) == \
        (_lib.FLAG_DONE, _lib.FLAG_BAD_ACTION, _lib.FLAG_STEPPED_DONE, _lib.FLAG_BAD_STATE)
    assert [d["POMDP_KIND_" + k] for k in ("ROCK", "TAG", "BATTLESHIP", "TIGER", "NETWORK")] == \
        [_lib.KIND_ROCK, _lib.KIND_TAG, _lib.KIND_BATTLESHIP, _lib.KIND_TIGER, _lib.KIND_NETWORK]
    assert [d["POMDP_COORD_" + k] for k in ("GET_INDEX", "GET_COORD", "IS_INSIDE", "ADD_MOVE", "L1", "TAG_GET_INDEX",
                                             "TAG_GET_COORD", "TAG_IS_INSIDE")] == list(range(8))


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.RockParams) == 24
    assert ctypes.sizeof(_lib.TagParams) == 16
    assert ctypes.sizeof(_lib.BattleshipParams) == 16
    assert ctypes.sizeof(_lib.TigerParams) == 8
    assert ctypes.sizeof(_lib.NetworkParams) == 32


def _bound(path):
    return _lib._bind(path)


def test_host_side_calls_of_the_product_library(product):
    """Parameter queries and table construction are host-only: they work without a GPU and
    agree with the host simulation byte for byte (same pomdp_host.h)."""
    P, H = _bound(build.OUT), _bound(build_hostsim())
    assert P.pomdp_abi_version() == H.pomdp_abi_version() == _lib.ABI_VERSION
    for board, rocks, stoch in [(7, 8, 0), (11, 11, 0), (15, 15, 0), (7, 7, 1), (4, 3, 0), (2, 1, 0)]:
        q = _lib.RockParams(board, rocks, stoch, 0, 0.8)
        words = P.pomdp_rock_state_words(ctypes.byref(q))
        assert words == (1 if rocks <= 11 else 2) == H.pomdp_rock_state_words(ctypes.byref(q))
        nb = P.pomdp_rock_table_bytes(ctypes.byref(q))
        assert nb == H.pomdp_rock_table_bytes(ctypes.byref(q)) and nb % 16 == 0 and nb > 400
        a, b = np.zeros(nb, np.uint8), np.ones(nb, np.uint8)
        assert P.pomdp_rock_build_table(ctypes.byref(q), a.ctypes.data) == 0
        assert H.pomdp_rock_build_table(ctypes.byref(q), b.ctypes.data) == 0
        assert np.array_equal(a, b)
    assert P.pomdp_belief_hist_bins(_lib.KIND_ROCK, 11, 1) == 11 + 256
    assert P.pomdp_belief_hist_bins(_lib.KIND_TAG, 0, 0) == 58
    assert P.pomdp_belief_hist_bins(_lib.KIND_BATTLESHIP, 100, 0) == 100
    assert P.pomdp_belief_hist_bins(_lib.KIND_TIGER, 0, 0) == 2
    assert P.pomdp_belief_hist_bins(_lib.KIND_NETWORK, 10, 0) == 10
    assert P.pomdp_belief_hist_bins(99, 0, 0) < 0


@pytest.mark.parametrize("which", ["product", "hostsim"])
def test_error_convention(which):
    """Bad arguments are rejected on the host BEFORE any launch: POMDP_E_BADARG / POMDP_E_ALIGN
    plus a thread-local message; no exception crosses the ABI (rock.py:101's assert, the
    reference's other constructor checks)."""
    L = _bound(build.OUT if which == "product" else build_hostsim())
    E_BADARG, E_ALIGN = -1, -2
    bad = _lib.RockParams(9, 9, 0, 0, 0.8)                      # not a key of rock.config
    assert L.pomdp_rock_state_words(ctypes.byref(bad)) == E_BADARG
    assert b"rock.config" in L.pomdp_last_error()
    bad = _lib.RockParams(7, 9, 0, 0, 0.8)                      # num_rocks not in config[7]['size']
    assert L.pomdp_rock_table_bytes(ctypes.byref(bad)) == E_BADARG
    assert L.pomdp_rock_build_table(ctypes.byref(_lib.RockParams(7, 8, 0, 0, .8)), None) == E_BADARG
    ok = _lib.RockParams(7, 8, 0, 0, 0.8)
    buf = np.zeros(64, np.int32)
    p = buf.ctypes.data
    # negative n, NULL pointers, misaligned pointers: refused before touching the device
    assert L.pomdp_rock_step(ctypes.byref(ok), p, p, p, p, p, p, p, -1, 0, 0, 0, None) == E_BADARG
    assert L.pomdp_rock_step(ctypes.byref(ok), p, None, p, p, p, p, p, 4, 0, 0, 0, None) == E_BADARG
    assert L.pomdp_rock_step(ctypes.byref(ok), p, p + 2, p, p, p, p, p, 4, 0, 0, 0, None) == E_ALIGN
    assert b"aligned" in L.pomdp_last_error()
    # n == 0 is a valid no-op everywhere (empty particle sets)
    assert L.pomdp_rock_step(ctypes.byref(ok), p, None, None, None, None, None, None, 0, 0, 0, 0, None) == 0
    assert L.pomdp_tag_step(ctypes.byref(_lib.TagParams(1, 0, .8)), None, None, None, None, None, None, None, 0, 0, 0, 0,
                            None) == 0
    assert L.pomdp_tag_step(ctypes.byref(_lib.TagParams(7, 0, .8)), p, p, p, p, p, p, p, 4, 0, 0, 0, None) == E_BADARG
    assert L.pomdp_tag_table_bytes() == 25472 and L.pomdp_tag_build_table(None) == E_BADARG
    # rollouts / policy: same conventions
    d8 = np.zeros(8, np.float64).ctypes.data
    assert L.pomdp_rock_rollout(ctypes.byref(ok), p, p, None, p, d8, p, p, 4, 0, 0, 0, -1, .95, None) == E_BADARG      # max_steps < 0
    assert L.pomdp_rock_rollout(ctypes.byref(ok), p, p, None, p, d8 + 4, p, p, 4, 0, 0, 0, 5, .95, None) == E_ALIGN    # float64 returns
    assert L.pomdp_rock_rollout(ctypes.byref(ok), p, None, None, None, None, None, None, 0, 0, 0, 0, 5, .95, None) == 0
    assert L.pomdp_tiger_policy(ctypes.byref(_lib.TigerParams(.85)), p, None, 4, 0, 0, 0, None) == E_BADARG
    assert L.pomdp_tiger_policy(ctypes.byref(_lib.TigerParams(.85)), p, p, 4, -4, 0, 0, None) == E_BADARG         # global_offset < 0
    assert L.pomdp_network_step(ctypes.byref(_lib.NetworkParams(31, 3, .1, .33, .95)), p, p, p, p, p, p, 4, 0, 0, 0,
                                None) == E_BADARG
    assert L.pomdp_network_step(ctypes.byref(_lib.NetworkParams(9, 3, .1, .33, .95)), p, p, p, p, p, p, 4, 0, 0, 0,
                                None) == E_BADARG        # network.py:155: n % 3 == 1
    assert L.pomdp_battleship_step(ctypes.byref(_lib.BattleshipParams(20, 20, 3, 0)), p, p, p, p, p, p, 4, None) == E_BADARG
    assert L.pomdp_battleship_reset(ctypes.byref(_lib.BattleshipParams(5, 5, 1, 0)), None, p, p, p, None, 4, 0, 0, 0,
                                    None) == E_BADARG
    assert L.pomdp_battleship_reset(ctypes.byref(_lib.BattleshipParams(5, 5, 3, 0)), p + 4, p, p, p, None, 4, 0, 0, 0,
                                    None) == E_ALIGN         # d_table must be 16-byte aligned
    assert L.pomdp_battleship_table_bytes(ctypes.byref(_lib.BattleshipParams(20, 20, 3, 0))) < 0
    assert L.pomdp_battleship_build_table(ctypes.byref(_lib.BattleshipParams(5, 5, 3, 0)), None) == E_BADARG
    assert L.pomdp_coord_op(99, 7, 7, p, None, p, 4, None) == E_BADARG
    assert L.pomdp_coord_op(_lib.COORD_L1, 7, 7, p, None, p, 4, None) == E_BADARG     # L1 needs b
    assert L.pomdp_belief_hist(99, 0, 0, p, 1, 4, p, None) == E_BADARG


def test_no_cpu_fallback_in_the_product():
    """The shipped package never imports oracle/ or the host simulation, and refuses CPU devices."""
    pkg = os.path.join(ROOT, "gym_pomdp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libpomdp_oracle" not in src and "c_oracle" not in src, f
    import gym_pomdp_b200 as gp
    _lib._inject_for_tests(None)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        gp.make("Tiger-v0", batch_size=4, device="cpu")
