from .agent import Agent
from .a2c import A2C
from .ddpg import DDPG
from .ppo import PPO
from .sac import SAC
from .td3 import TD3

__all__ = [Agent, A2C, DDPG, PPO, SAC, TD3]
