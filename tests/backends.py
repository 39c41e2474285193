"""Test helpers: pick the library the env classes talk to.

* ``cuda``    -- the product: gym_pomdp_b200/csrc/libpomdp_b200.so on a real GPU (``-m gpu``).
* ``hostsim`` -- tests/hostsim/libpomdp_hostsim.so: the same per-env functors compiled with
  g++ (TEST VEHICLE; lets the CPU suite check packed-state logic and the Python host layer).
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_SO = os.path.join(HOSTSIM_DIR, "libpomdp_hostsim.so")
GOLDEN_SEED = 0x5EED


def build_hostsim():
    if os.environ.get("POMDP_HOSTSIM_SO"):          # an instrumented build (tests/test_hostsim_ubsan.py)
        return os.environ["POMDP_HOSTSIM_SO"]
    src = os.path.join(HOSTSIM_DIR, "pomdp_hostsim.cpp")
    deps = [src, os.path.join(ROOT, "gym_pomdp_b200", "csrc", "pomdp_core.h"), os.path.join(ROOT, "gym_pomdp_b200", "csrc", "pomdp_envs.h"),
            os.path.join(ROOT, "gym_pomdp_b200", "csrc", "pomdp_host.h"), os.path.join(ROOT, "include", "pomdp_b200.h")]
    if os.path.exists(HOSTSIM_SO) and all(os.path.getmtime(d) <= os.path.getmtime(HOSTSIM_SO) for d in deps):
        return HOSTSIM_SO
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", HOSTSIM_SO, src],
                   check=True)
    return HOSTSIM_SO


BACKENDS = [pytest.param("hostsim", id="hostsim"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request):
    """Yields the torch device string to build envs on, with the right library bound."""
    from gym_pomdp_b200 import _lib
    if request.param == "hostsim":
        _lib._inject_for_tests(build_hostsim())
        yield "cpu"
        _lib._inject_for_tests(None)
    else:
        import torch
        _lib._inject_for_tests(None)
        assert torch.cuda.is_available(), "-m gpu tests need a GPU"
        assert not _lib.is_hostsim()
        yield "cuda:0"


def philox_unmodified(draws, stream, domain, n_slots=None):
    """Rows of a fixture's draw table that are still the plain Philox words (the generator
    overwrote a few with Bernoulli-boundary values the kernels cannot be fed directly)."""
    from oracle import philox
    n_slots = draws.shape[1] if n_slots is None else n_slots
    ref = philox.draw_slots(GOLDEN_SEED, np.arange(draws.shape[0]), stream, domain, n_slots)
    return (ref == draws[:, :n_slots]).all(axis=1)
