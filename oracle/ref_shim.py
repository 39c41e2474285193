"""Import the UNMODIFIED reference (d3sm0/gym_pomdp @ /root/reference) in this container
and drive its random draws from a scripted stream.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` (which writes the committed
fixtures under tests/golden/) and by the optional local tests that re-check the oracle
against the live reference.  /root/reference does not exist on the GPU box, so nothing
that runs there (``-m gpu`` tests, smoke(), bench.py) may import this module's
``load_reference``.

Why a stub: the reference needs ``gym`` and ``pygame`` at import time (rock.py:5-9,
tiger.py:3-7, gui.py:13) and neither is installed here (no network).  It uses only
``gym.Env``/``gym.core.Env`` as a base class, ``gym.spaces.Discrete`` (``n``,
``contains``, ``sample``) and ``gym.envs.registration.register``; ``pygame`` only through
``pygame.Color`` in gui.py:13.  The stub below provides exactly those names and no
hot-path arithmetic.  ``Discrete.sample`` is the one semantic choice (uniform over
``range(n)``, what gym does); it only affects Tiger's state draw (tiger.py:64,119,123).

Coupling rule (what makes element-wise stochastic parity possible).  With a 32-bit draw
word ``r`` and ``u = r / 2**32`` (exact in a double):

    np.random.binomial(1, p)  ->  1 if u < p else 0
    np.random.uniform(a, b)   ->  a + (b - a) * u
    np.random.randint(n)      ->  (r * n) >> 32          (= floor(u * n))
    np.random.choice(seq)     ->  seq[(r * len(seq)) >> 32]
    Discrete(n).sample()      ->  (r * n) >> 32          (separate queue: gym's own RNG)

The kernels and the C/Python oracles apply the same integer rules to the same Philox
words, so a mismatch anywhere is a semantic difference, not RNG noise.
"""
import contextlib
import importlib
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
TWO32 = float(2 ** 32)


class ScriptedDraws:
    """FIFO of uint32 draw words consumed by the patched numpy / gym entry points."""

    def __init__(self):
        self.np_queue = []
        self.gym_queue = []
        self.consumed = 0

    def feed(self, words):
        self.np_queue.extend(int(w) for w in words)

    def feed_gym(self, words):
        self.gym_queue.extend(int(w) for w in words)

    def clear(self):
        self.np_queue.clear()
        self.gym_queue.clear()

    def _pop(self):
        if not self.np_queue:
            raise RuntimeError("reference consumed more numpy draws than were scripted")
        self.consumed += 1
        return self.np_queue.pop(0)

    def _pop_gym(self):
        if not self.gym_queue:
            raise RuntimeError("reference consumed more gym draws than were scripted")
        return self.gym_queue.pop(0)

    # -- the four numpy entry points the reference uses (SURVEY.md §4 tier 2) --
    def binomial(self, n, p=None, size=None):
        assert n == 1 and size is None
        return 1 if (self._pop() / TWO32) < p else 0

    def uniform(self, low=0.0, high=1.0, size=None):
        assert size is None
        return low + (high - low) * (self._pop() / TWO32)

    def randint(self, low, high=None, size=None):
        assert size is None
        if high is None:
            low, high = 0, low
        return low + ((self._pop() * (high - low)) >> 32)

    def choice(self, a, size=None, replace=True, p=None):
        assert size is None and p is None
        return a[(self._pop() * len(a)) >> 32]


_DRAWS = ScriptedDraws()


def draws():
    return _DRAWS


def _install_stubs():
    if "gym" in sys.modules and getattr(sys.modules["gym"], "__pomdp_stub__", False):
        return
    gym = types.ModuleType("gym")
    gym.__pomdp_stub__ = True

    class Env(object):
        metadata = {}

    class Discrete(object):
        def __init__(self, n):
            self.n = int(n)

        def contains(self, x):
            try:
                return int(x) == x and 0 <= int(x) < self.n
            except (TypeError, ValueError):
                return False

        def sample(self):
            return (_DRAWS._pop_gym() * self.n) >> 32

    registry = {}

    def register(id, entry_point=None, **kwargs):
        registry[id] = (entry_point, kwargs)

    core = types.ModuleType("gym.core")
    core.Env = Env
    spaces = types.ModuleType("gym.spaces")
    spaces.Discrete = Discrete
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = register
    registration.registry = registry
    envs.registration = registration
    gym.Env, gym.core, gym.spaces, gym.envs = Env, core, spaces, envs
    sys.modules.update({"gym": gym, "gym.core": core, "gym.spaces": spaces,
                        "gym.envs": envs, "gym.envs.registration": registration})

    pygame = types.ModuleType("pygame")
    pygame.__pomdp_stub__ = True
    pygame.Color = lambda *a, **k: (128, 128, 128)
    sys.modules["pygame"] = pygame


def reference_root():
    """Where the unmodified reference package lies: /root/reference in the build container, else the copy that
    ``make -C oracle _ref`` put under oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot)."""
    import os
    for root in (REFERENCE_ROOT, os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")):
        if os.path.isdir(os.path.join(root, "gym_pomdp", "envs")):
            return root
    return None


def load_reference():
    """Returns the reference's ``gym_pomdp.envs`` package (imported, not copied)."""
    _install_stubs()
    root = reference_root()
    if root is None:
        raise ImportError("the reference is neither at %s nor under oracle/_ref (make -C oracle _ref)" % REFERENCE_ROOT)
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("gym_pomdp.envs")


def reference_available():
    return reference_root() is not None


@contextlib.contextmanager
def scripted_numpy():
    """Patch the four np.random entry points for the duration of the block."""
    saved = {k: getattr(np.random, k) for k in ("binomial", "uniform", "randint", "choice")}
    np.random.binomial = _DRAWS.binomial
    np.random.uniform = _DRAWS.uniform
    np.random.randint = _DRAWS.randint
    np.random.choice = _DRAWS.choice
    try:
        yield _DRAWS
    finally:
        for k, v in saved.items():
            setattr(np.random, k, v)
        _DRAWS.clear()
