from .battleship import BattleShipEnv
from .network import NetworkEnv
from .rock import RockEnv, StochasticRockEnv
from .tag import TagEnv
from .tiger import TigerEnv

__all__ = ["BattleShipEnv", "NetworkEnv", "RockEnv", "StochasticRockEnv", "TagEnv", "TigerEnv"]
