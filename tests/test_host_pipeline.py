"""Host-buffer entry point (pomdp_step_packed_host, include/pomdp_b200.h): what a numpy-holding caller of the
reference does with a whole particle set -- (state, action) in host memory in, (next_state, result) in host memory
out -- must equal one packed step over the same batch on device-resident tensors, for any chunking."""
import ctypes

import numpy as np
import pytest
import torch

import gym_pomdp_b200 as gp
from gym_pomdp_b200 import _lib

from backends import backend  # noqa: F401
from test_edge_cases import make_all, random_inputs

NAMES = ["rock", "rock15", "srock", "tag", "tag3", "tiger", "network"]


def host(t, dev):
    t = t.cpu().contiguous()
    return t.pin_memory() if dev != "cpu" else t


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("chunk,n_slots", [(1024, 3), (4, 1), (8192, 2), (1 << 16, 8)])
def test_host_pipeline_equals_device_step(backend, name, chunk, n_slots):
    n = 10007 if chunk > 4 else 203                                       # ragged: the last chunk is short
    env = make_all(backend, n, global_offset=4 * 123)[name]
    rs = np.random.RandomState(3)
    state, action = random_inputs(env, name, n, rs, backend)
    want_state, want_result = env.simulate(state, action, step_ctr=17, packed=True)
    hs, ha = host(state, backend), host(action, backend)
    out = (host(torch.zeros_like(state), backend), host(torch.zeros_like(action), backend))
    got = env.simulate_host(hs, ha, out, step_ctr=17, packed=True, pipeline="c", chunk=chunk, n_streams=n_slots)
    assert torch.equal(got[0], want_state.cpu()) and torch.equal(got[1], want_result.cpu())
    # the same pipe again (slots are reused), then the torch-driven pipeline
    got = env.simulate_host(hs, ha, out, step_ctr=18, packed=True, pipeline="c", chunk=chunk, n_streams=n_slots)
    want2 = env.simulate(state, action, step_ctr=18, packed=True)
    assert torch.equal(got[0], want2[0].cpu()) and torch.equal(got[1], want2[1].cpu())
    if backend != "cpu":
        out2 = (host(torch.zeros_like(state), backend), host(torch.zeros_like(action), backend))
        py = env.simulate_host(hs, ha, out2, step_ctr=17, packed=True, pipeline="python", chunk=max(chunk, 1024))
        assert torch.equal(py[0], want_state.cpu()) and torch.equal(py[1], want_result.cpu())
    env.close()


def test_host_pipeline_empty_and_default(backend):
    env = make_all(backend, 64)["rock"]
    z = torch.zeros(0, dtype=torch.int32)
    out = env.simulate_host(z, z.clone(), (z.clone(), z.clone()), step_ctr=1, packed=True, pipeline="c")
    assert out[0].numel() == 0 and out[1].numel() == 0
    env.close()


def test_host_pipeline_argument_errors(backend):
    L = _lib.lib()
    h = ctypes.c_void_p()
    E, E_ALIGN = -1, -2                                                     # POMDP_E_BADARG, POMDP_E_ALIGN
    assert L.pomdp_host_pipe_create(1, 1000, 3, None) == E
    assert L.pomdp_host_pipe_create(1, 1001, 3, ctypes.byref(h)) == E             # chunk not a multiple of 4
    assert L.pomdp_host_pipe_create(1, 1024, 0, ctypes.byref(h)) == E
    assert L.pomdp_host_pipe_create(1, 1024, 9, ctypes.byref(h)) == E
    assert L.pomdp_host_pipe_create(0, 1024, 3, ctypes.byref(h)) == E
    assert L.pomdp_host_pipe_destroy(None) == E
    envs = make_all(backend, 16)
    ctxt = torch.cuda.device(0) if backend != "cpu" else None
    if ctxt:
        ctxt.__enter__()
    try:
        assert L.pomdp_host_pipe_create(1, 1024, 2, ctypes.byref(h)) == 0
        rock, rock15, tag = envs["rock"], envs["rock15"], envs["tag"]
        buf = torch.zeros(64, dtype=torch.int32)
        p = buf.data_ptr()
        args = (p, p, p, p, 16, 0, 1, 1)
        assert L.pomdp_step_packed_host(None, _lib.KIND_ROCK, ctypes.addressof(rock._params), _lib.ptr(rock._table), *args) == E
        assert L.pomdp_step_packed_host(h, _lib.KIND_BATTLESHIP, ctypes.addressof(envs["ship"]._params), None, *args) == E
        assert L.pomdp_step_packed_host(h, 99, ctypes.addressof(rock._params), None, *args) == E
        assert L.pomdp_step_packed_host(h, _lib.KIND_ROCK, None, None, *args) == E
        # a pipe made for one state word cannot carry Rock(15,15)'s two
        assert L.pomdp_step_packed_host(h, _lib.KIND_ROCK, ctypes.addressof(rock15._params), _lib.ptr(rock15._table), *args) == E
        assert b"state words" in L.pomdp_last_error()
        assert L.pomdp_step_packed_host(h, _lib.KIND_TAG, ctypes.addressof(tag._params), _lib.ptr(tag._table), None, p, p, p, 16, 0, 1, 1) == E
        assert L.pomdp_step_packed_host(h, _lib.KIND_TAG, ctypes.addressof(tag._params), _lib.ptr(tag._table), p, p, p, p, -1, 0, 1, 1) == E
        assert L.pomdp_step_packed_host(h, _lib.KIND_TAG, ctypes.addressof(tag._params), _lib.ptr(tag._table), p + 2, p, p, p, 4, 0, 1, 1) == E_ALIGN
        assert L.pomdp_step_packed_host(h, _lib.KIND_TAG, ctypes.addressof(tag._params), _lib.ptr(tag._table), p, p, p, p, 0, 0, 1, 1) == 0
        assert L.pomdp_host_pipe_destroy(h) == 0
    finally:
        if ctxt:
            ctxt.__exit__(None, None, None)
    with pytest.raises(ValueError):
        rock.simulate_host(buf, buf, (buf, buf, buf, buf), packed=False, pipeline="c")
    with pytest.raises(ValueError):
        rock.simulate_host(buf.long(), buf, (buf, buf), packed=True, pipeline="c")
