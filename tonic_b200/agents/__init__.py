from .agent import Agent

__all__ = [Agent]
