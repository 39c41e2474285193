"""The reference's ``gym.make()`` ids (gym_pomdp/__init__.py:7-46), kept.

In scope: ``Tiger-v0``, ``Tag-v0``, ``Battleship-v0``, ``Rock-v0``, ``StochasticRock-v0``,
``Network-v0``.  (The reference registers ``StochasticRock-v0`` with a misspelt entry point
-- ``StochasticRockEvn``, __init__.py:35 -- so its own ``gym.make`` of that id fails; it is
registered correctly here.)  ``Pocman-v0`` and ``Test-v0`` are out of scope.

``make(id, **kwargs)`` works without gym.  If ``gym`` or ``gymnasium`` is importable the ids are also registered there
(neither package is in the build image; tests/test_gym_surface.py drives this against stubs of both APIs):

* old gym (< 0.26), the API the reference is written against -- ``reset() -> ob``, ``step -> (ob, rw, done, info)`` -- gets
  the env classes themselves (they subclass ``gym.Env`` there);
* gym >= 0.26 and gymnasium call ``reset(seed=..., options=...)`` and expect ``(ob, info)`` / a 5-tuple from ``step``, and
  wrap what ``make`` returns in checkers that insist on it: they get ``NewApiAdapter`` around the same env (``.env`` is the
  old-API object with all the planner hooks; unknown attributes are forwarded to it), registered with the passive
  checker off.
"""
import functools
import importlib

ENTRY_POINTS = {
    "Tiger-v0": "gym_pomdp_b200.envs:TigerEnv",
    "Tag-v0": "gym_pomdp_b200.envs:TagEnv",
    "Battleship-v0": "gym_pomdp_b200.envs:BattleShipEnv",
    "Rock-v0": "gym_pomdp_b200.envs:RockEnv",
    "StochasticRock-v0": "gym_pomdp_b200.envs:StochasticRockEnv",
    "Network-v0": "gym_pomdp_b200.envs:NetworkEnv",
}

registry = dict(ENTRY_POINTS)


def _load(entry_point):
    mod, _, name = entry_point.partition(":")
    return getattr(importlib.import_module(mod), name)


def make(id, **kwargs):
    """``gym.make`` for the ids above; extra kwargs go to the env constructor, e.g.
    ``make("Rock-v0", board_size=11, num_rocks=11, batch_size=1 << 22, device="cuda:0")``."""
    if id not in registry:
        raise KeyError("unknown environment id %r (known: %s)" % (id, ", ".join(sorted(registry))))
    return _load(registry[id])(**kwargs)


def _version(mod):
    out = []
    for part in str(getattr(mod, "__version__", "0")).split(".")[:2]:
        digits = "".join(ch for ch in part if ch.isdigit())
        out.append(int(digits) if digits else 0)
    return tuple(out)


def uses_new_api(pkg, mod=None):
    """gymnasium, and gym from 0.26 on: ``reset(seed, options) -> (ob, info)``; ``step`` returns five values."""
    mod = importlib.import_module(pkg) if mod is None else mod
    return pkg == "gymnasium" or _version(mod) >= (0, 26)


@functools.lru_cache(maxsize=None)
def _adapter_class(pkg):
    base = importlib.import_module(pkg).Env

    class NewApiAdapter(base):
        """The old-gym env (the reference's protocol) behind the reset(seed, options) / five-value step protocol."""
        metadata = {"render_modes": ["ansi"], "render.modes": ["ansi"]}

        def __init__(self, env):
            self.env = env
            self.action_space = env.action_space
            self.observation_space = env.observation_space

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self.env.seed(seed)
            mask = (options or {}).get("mask") if isinstance(options, dict) else None
            return self.env.reset(mask) if mask is not None else self.env.reset(), {}

        def step(self, action):
            ob, rw, done, info = self.env.step(action)
            truncated = False if isinstance(done, bool) else done & False        # batched mode: a tensor of the same shape
            return ob, rw, done, truncated, info

        def render(self, *a, **k):
            return self.env.render(*a, **k)

        def close(self):
            return self.env.close()

        def __getattr__(self, name):                                           # planner hooks: _set_state, _generate_legal, ...
            if name == "env":
                raise AttributeError(name)
            return getattr(self.env, name)

    return NewApiAdapter


def make_new_api(env_id, pkg="gymnasium", **kwargs):
    """What gym >= 0.26 / gymnasium get from ``gym.make(env_id, **kwargs)``."""
    kwargs.pop("render_mode", None)
    return _adapter_class(pkg)(make(env_id, **kwargs))


def register_with_gym():
    """Best effort; returns the name of the package the ids were registered with, or None."""
    for pkg in ("gym", "gymnasium"):
        try:
            mod = importlib.import_module(pkg)
            reg = importlib.import_module(pkg + ".envs.registration")
        except Exception:  # noqa: BLE001
            continue
        new_api = uses_new_api(pkg, mod)
        for env_id, entry in ENTRY_POINTS.items():
            try:
                if new_api:
                    reg.register(id=env_id, entry_point=functools.partial(make_new_api, env_id, pkg), disable_env_checker=True)
                else:
                    reg.register(id=env_id, entry_point=entry)
            except Exception:  # noqa: BLE001 - already registered
                pass
        return pkg
    return None
