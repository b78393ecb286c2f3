from . import agents, models, normalizers, updaters

__all__ = [agents, models, normalizers, updaters]
