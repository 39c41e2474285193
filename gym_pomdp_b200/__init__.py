"""gym_pomdp_b200 -- batched, B200-native step()/reset() for the gym_pomdp environments.

    import gym_pomdp_b200 as gp
    env = gp.make("Rock-v0", board_size=11, num_rocks=11, batch_size=1 << 22, device="cuda:0")
    obs = env.reset()
    obs, reward, done, info = env.step(actions)        # one sm_100a kernel launch

Importing the package does not load the CUDA library; constructing an environment does, and
fails loudly if it has not been built (``python -m gym_pomdp_b200.build``).
"""
from .registration import ENTRY_POINTS, make, register_with_gym, registry  # noqa: F401

__version__ = "0.1.0"

GYM_BACKEND = register_with_gym()
