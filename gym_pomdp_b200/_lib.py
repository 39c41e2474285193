"""ctypes binding of libpomdp_b200.so (the C ABI declared in include/pomdp_b200.h).

The product has exactly one backend: the CUDA library built in-tree by
``gym_pomdp_b200.build`` (``python -m gym_pomdp_b200.build``).  If it is missing, loading
fails loudly -- there is no CPU fallback and nothing here imports oracle/.

``_inject_for_tests`` exists for the CPU-only test-suite: it swaps in
tests/hostsim/libpomdp_hostsim.so (the same per-env functors compiled with g++, same
symbols, host pointers) so the Python host logic can be exercised without a GPU.  It is
never called from product code.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int32, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# POMDP_B200_LIB lets kernel-tuning experiments (scripts/exp_variants.sh) point at another build of the SAME library
LIB_PATH = os.environ.get("POMDP_B200_LIB") or os.path.join(_HERE, "csrc", "libpomdp_b200.so")

ABI_VERSION = 15
FLAG_DONE = 1
FLAG_BAD_ACTION = 2
FLAG_STEPPED_DONE = 4
FLAG_BAD_STATE = 8
FLAG_ERRORS = FLAG_BAD_ACTION | FLAG_STEPPED_DONE | FLAG_BAD_STATE

KIND_ROCK, KIND_TAG, KIND_BATTLESHIP, KIND_TIGER, KIND_NETWORK = range(5)

COORD_GET_INDEX, COORD_GET_COORD, COORD_IS_INSIDE, COORD_ADD_MOVE, COORD_L1, \
    COORD_TAG_GET_INDEX, COORD_TAG_GET_COORD, COORD_TAG_IS_INSIDE = range(8)


class RockParams(ctypes.Structure):
    _fields_ = [("board_size", c_int32), ("num_rocks", c_int32), ("stochastic", c_int32),
                ("reserved", c_int32), ("p_move", c_double)]


class RockHeuristicPlanes(ctypes.Structure):
    """PomdpRockHeuristicPlanes: device pointers (0 = the fresh value / an empty history)"""
    _fields_ = [("count", c_void_p), ("measured", c_void_p), ("lkv", c_void_p), ("lkw", c_void_p), ("prob_valuable", c_void_p),
                ("check_totals", c_void_p), ("prev_obs", c_void_p), ("scratch", c_void_p)]


class TagParams(ctypes.Structure):
    _fields_ = [("num_opponents", c_int32), ("reserved", c_int32), ("move_prob", c_double)]


class BattleshipParams(ctypes.Structure):
    _fields_ = [("x_size", c_int32), ("y_size", c_int32), ("max_len", c_int32), ("reserved", c_int32)]


class TigerParams(ctypes.Structure):
    _fields_ = [("listen_prob", c_double)]


class NetworkParams(ctypes.Structure):
    _fields_ = [("n_machines", c_int32), ("problem_type", c_int32), ("p", c_double), ("q", c_double),
                ("p_ob", c_double)]


class HistSink(ctypes.Structure):
    """PomdpHistSink: where a step kernel with the histogram epilogue (pomdp_E_step_hist) hands its counts"""
    _fields_ = [("scratch", c_void_p), ("d_peer_bufs", c_void_p), ("world", c_int32), ("rank", c_int32), ("wait", c_int32),
                ("pad_", c_int32), ("hist_out", c_void_p)]


_P = c_void_p  # device (or, under hostsim, host) array pointers travel as raw addresses
_STEP_TAIL = [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64, c_uint32, c_void_p]
_STEP_HIST_TAIL = [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64, c_uint32, POINTER(HistSink), c_void_p]
_RESET_TAIL = [_P, _P, _P, c_int64, c_int64, c_uint64, c_uint32, c_void_p]

_POLICY_TAIL = [_P, _P, c_int64, c_int64, c_uint64, c_uint32, c_void_p]
_ROLLOUT_TAIL = [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64, c_uint32, c_int32, c_double, c_void_p]

_STEPP_TAIL = [_P, _P, _P, _P, c_int64, c_int64, c_uint64, c_uint32, c_void_p]

_PROTOTYPES = {
    "pomdp_abi_version": (c_int32, []),
    "pomdp_last_error": (c_char_p, []),
    "pomdp_rock_state_words": (c_int32, [POINTER(RockParams)]),
    "pomdp_rock_table_bytes": (c_int64, [POINTER(RockParams)]),
    "pomdp_rock_build_table": (c_int32, [POINTER(RockParams), c_void_p]),
    "pomdp_rock_step": (c_int32, [POINTER(RockParams), _P] + _STEP_TAIL),
    "pomdp_rock_reset": (c_int32, [POINTER(RockParams), _P] + _RESET_TAIL),
    "pomdp_tag_table_bytes": (c_int64, []),
    "pomdp_tag_build_table": (c_int32, [c_void_p]),
    "pomdp_tag_step": (c_int32, [POINTER(TagParams), _P] + _STEP_TAIL),
    "pomdp_tag_reset": (c_int32, [POINTER(TagParams)] + _RESET_TAIL),
    "pomdp_battleship_step": (c_int32, [POINTER(BattleshipParams), _P, _P, _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_battleship_table_bytes": (c_int64, [POINTER(BattleshipParams)]),
    "pomdp_battleship_build_table": (c_int32, [POINTER(BattleshipParams), c_void_p]),
    "pomdp_battleship_reset": (c_int32, [POINTER(BattleshipParams), _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64,
                                         c_uint32, c_void_p]),
    "pomdp_battleship_reset_warpscan": (c_int32, [POINTER(BattleshipParams), _P, _P, _P, _P, c_int64, c_int64, c_uint64,
                                                  c_uint32, c_void_p]),
    "pomdp_battleship_reset_rejection": (c_int32, [POINTER(BattleshipParams), _P, _P, _P, _P, c_int64, c_int64,
                                                   c_uint64, c_uint32, c_void_p]),
    "pomdp_tiger_step": (c_int32, [POINTER(TigerParams)] + _STEP_TAIL),
    "pomdp_tiger_reset": (c_int32, [POINTER(TigerParams)] + _RESET_TAIL),
    "pomdp_network_step": (c_int32, [POINTER(NetworkParams)] + _STEP_TAIL),
    "pomdp_network_reset": (c_int32, [POINTER(NetworkParams), _P, _P, _P, c_int64, c_void_p]),
    "pomdp_rock_step_packed": (c_int32, [POINTER(RockParams), _P] + _STEPP_TAIL),
    "pomdp_tag_step_packed": (c_int32, [POINTER(TagParams), _P] + _STEPP_TAIL),
    "pomdp_tiger_step_packed": (c_int32, [POINTER(TigerParams)] + _STEPP_TAIL),
    "pomdp_network_step_packed": (c_int32, [POINTER(NetworkParams)] + _STEPP_TAIL),
    "pomdp_host_pipe_create": (c_int32, [c_int32, c_int64, c_int32, POINTER(c_void_p)]),
    "pomdp_host_pipe_destroy": (c_int32, [c_void_p]),
    "pomdp_step_packed_host": (c_int32, [c_void_p, c_int32, c_void_p, _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64, c_uint32]),
    "pomdp_rock_policy": (c_int32, [POINTER(RockParams), _P] + _POLICY_TAIL),
    "pomdp_rock_rollout": (c_int32, [POINTER(RockParams), _P] + _ROLLOUT_TAIL),
    "pomdp_tag_policy": (c_int32, [POINTER(TagParams), _P] + _POLICY_TAIL),
    "pomdp_tag_rollout": (c_int32, [POINTER(TagParams), _P] + _ROLLOUT_TAIL),
    "pomdp_battleship_policy": (c_int32, [POINTER(BattleshipParams)] + _POLICY_TAIL),
    "pomdp_battleship_rollout": (c_int32, [POINTER(BattleshipParams)] + _ROLLOUT_TAIL),
    "pomdp_tiger_policy": (c_int32, [POINTER(TigerParams)] + _POLICY_TAIL),
    "pomdp_tiger_rollout": (c_int32, [POINTER(TigerParams)] + _ROLLOUT_TAIL),
    "pomdp_network_policy": (c_int32, [POINTER(NetworkParams)] + _POLICY_TAIL),
    "pomdp_network_rollout": (c_int32, [POINTER(NetworkParams)] + _ROLLOUT_TAIL),
    "pomdp_rock_obs_prob": (c_int32, [POINTER(RockParams), _P, _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_rock_legal_mask": (c_int32, [POINTER(RockParams), _P, _P, _P, c_int64, c_void_p]),
    "pomdp_tag_obs_prob": (c_int32, [POINTER(TagParams), _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_tag_legal_mask": (c_int32, [POINTER(TagParams), _P, _P, c_int64, c_void_p]),
    "pomdp_battleship_obs_prob": (c_int32, [POINTER(BattleshipParams), _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_battleship_legal_mask": (c_int32, [POINTER(BattleshipParams), _P, _P, c_int64, c_void_p]),
    "pomdp_tiger_obs_prob": (c_int32, [POINTER(TigerParams), _P, _P, _P, _P, c_int64, c_double, c_void_p]),
    "pomdp_tiger_legal_mask": (c_int32, [POINTER(TigerParams), _P, _P, c_int64, c_void_p]),
    "pomdp_network_obs_prob": (c_int32, [POINTER(NetworkParams), _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_network_legal_mask": (c_int32, [POINTER(NetworkParams), _P, _P, c_int64, c_void_p]),
    "pomdp_rock_belief_update": (c_int32, [POINTER(RockParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_rock_legal_list": (c_int32, [POINTER(RockParams), _P, _P, _P, c_int64, c_void_p]),
    "pomdp_rock_history_update": (c_int32, [POINTER(RockParams), _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_rock_preferred_mask": (c_int32, [POINTER(RockParams), _P, _P, _P, _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_rock_policy_preferred": (c_int32, [POINTER(RockParams), _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64,
                                              c_uint32, c_void_p]),
    "pomdp_rock_rollout_preferred": (c_int32, [POINTER(RockParams), _P, _P, _P, POINTER(RockHeuristicPlanes), _P, _P, _P, _P,
                                               c_int64, c_int64, c_uint64, c_uint32, c_int32, c_double, c_int32, c_void_p]),
    "pomdp_tag_preferred_mask": (c_int32, [POINTER(TagParams), _P, _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_tag_policy_preferred": (c_int32, [POINTER(TagParams), _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64, c_uint32,
                                             c_void_p]),
    "pomdp_tag_rollout_preferred": (c_int32, [POINTER(TagParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64,
                                              c_uint32, c_int32, c_double, c_void_p]),
    "pomdp_stream_probe": (c_int32, [_P, _P, _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_stream_probe_words": (c_int32, [c_int32, _P, _P, _P, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_coord_op": (c_int32, [c_int32, c_int32, c_int32, _P, _P, _P, c_int64, c_void_p]),
    "pomdp_belief_hist_bins": (c_int32, [c_int32, c_int32, c_int32]),
    "pomdp_belief_hist": (c_int32, [c_int32, c_int32, c_int32, _P, c_int32, c_int64, _P, c_void_p]),
    "pomdp_rock_step_hist": (c_int32, [POINTER(RockParams), _P] + _STEP_HIST_TAIL),
    "pomdp_tag_step_hist": (c_int32, [POINTER(TagParams), _P] + _STEP_HIST_TAIL),
    "pomdp_tiger_step_hist": (c_int32, [POINTER(TigerParams)] + _STEP_HIST_TAIL),
    "pomdp_network_step_hist": (c_int32, [POINTER(NetworkParams)] + _STEP_HIST_TAIL),
    "pomdp_belief_hist_once": (c_int32, [c_int32, c_int32, c_int32, _P, c_int32, c_int64, _P, _P, c_void_p]),
    "pomdp_belief_hist_allreduce": (c_int32, [c_int32, c_int32, c_int32, _P, c_int32, c_int64, _P, _P, c_int32, c_int32, c_int32,
                                              _P, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None
_is_hostsim = False


def _bind(path):
    lib = ctypes.CDLL(path)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError = a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


def lib():
    """The loaded CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m gym_pomdp_b200.build` "
                "(nvcc, sm_100a).  gym_pomdp_b200 has no CPU fallback.")
        _lib = _bind(LIB_PATH)
        got = _lib.pomdp_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libpomdp_b200.so ABI version {got}, expected {ABI_VERSION}")
    return _lib


def is_hostsim():
    return _is_hostsim


def _inject_for_tests(path):
    """TESTS ONLY: bind the g++-compiled host simulation of the kernels instead.  Refused outside a pytest run (pytest
    sets PYTEST_CURRENT_TEST for the duration of every test; worker processes a test spawns inherit it), so no product
    process can end up on a CPU library."""
    global _lib, _is_hostsim
    if path is None:
        _lib, _is_hostsim = None, False
        return
    if "PYTEST_CURRENT_TEST" not in os.environ:
        raise RuntimeError("_inject_for_tests is only available inside a pytest run; gym_pomdp_b200 has no CPU path")
    cand = _bind(path)
    assert hasattr(cand, "pomdp_is_hostsim"), "refusing to inject a library that is not the hostsim"
    _lib, _is_hostsim = cand, True


def check(rc, what=""):
    if rc != 0:
        msg = lib().pomdp_last_error()
        raise RuntimeError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t):
    """Raw address of a torch tensor's first element (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_handle(device):
    """cudaStream_t of torch's current stream on `device` (0 for host tensors under hostsim)."""
    import torch
    if device.type != "cuda":
        return None
    return torch.cuda.current_stream(device).cuda_stream
