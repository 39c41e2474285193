"""``Discrete`` space with the three members the reference uses from ``gym.spaces.Discrete``
(``n``, ``contains``, ``sample``; e.g. rock.py:113-114, rock.py:125, tiger.py:64).

If ``gym`` (or ``gymnasium``) is importable its own ``Discrete`` is used so that spaces
compare equal to what callers expect; neither is installed in the build image, so the
small stand-in below is what normally runs.
"""
import numpy as np

try:  # pragma: no cover - gym is absent in the build image
    from gym.spaces import Discrete as _GymDiscrete
except Exception:  # noqa: BLE001
    try:
        from gymnasium.spaces import Discrete as _GymDiscrete
    except Exception:  # noqa: BLE001
        _GymDiscrete = None


class _Discrete(object):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def contains(self, x):
        try:
            return int(x) == x and 0 <= int(x) < self.n
        except (TypeError, ValueError):
            return False

    __contains__ = contains

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return "Discrete(%d)" % self.n

    def __eq__(self, other):
        return hasattr(other, "n") and self.n == other.n


Discrete = _GymDiscrete if _GymDiscrete is not None else _Discrete
