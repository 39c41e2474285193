"""The reference's ``gym.make()`` ids (gym_pomdp/__init__.py:7-46), kept.

In scope: ``Tiger-v0``, ``Tag-v0``, ``Battleship-v0``, ``Rock-v0``, ``StochasticRock-v0``,
``Network-v0``.  (The reference registers ``StochasticRock-v0`` with a misspelt entry point
-- ``StochasticRockEvn``, __init__.py:35 -- so its own ``gym.make`` of that id fails; it is
registered correctly here.)  ``Pocman-v0`` and ``Test-v0`` are out of scope.

``make(id, **kwargs)`` works without gym.  If ``gym`` or ``gymnasium`` is importable the
ids are also registered there, pointing at the same classes, so ``gym.make("Rock-v0")``
keeps working for existing callers (neither package is in the build image).
"""
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
    import importlib
    mod, _, name = entry_point.partition(":")
    return getattr(importlib.import_module(mod), name)


def make(id, **kwargs):
    """``gym.make`` for the ids above; extra kwargs go to the env constructor, e.g.
    ``make("Rock-v0", board_size=11, num_rocks=11, batch_size=1 << 22, device="cuda:0")``."""
    if id not in registry:
        raise KeyError("unknown environment id %r (known: %s)" % (id, ", ".join(sorted(registry))))
    return _load(registry[id])(**kwargs)


def register_with_gym():
    """Best effort; returns the name of the package the ids were registered with, or None."""
    for pkg in ("gym", "gymnasium"):
        try:
            import importlib
            reg = importlib.import_module(pkg + ".envs.registration")
        except Exception:  # noqa: BLE001
            continue
        for env_id, entry in ENTRY_POINTS.items():
            try:
                reg.register(id=env_id, entry_point=entry)
            except Exception:  # noqa: BLE001 - already registered
                pass
        return pkg
    return None
