from .agent import Agent
from .a2c import A2C
from .ppo import PPO

__all__ = [Agent, A2C, PPO]
