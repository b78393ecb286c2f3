"""tonic_b200: B200-native backend for Tonic's data-parallel hot path."""

from . import agents, config, environments, explorations, replays  # noqa: E402,F401
from .utils import logger  # noqa: E402,F401
from .utils.trainer import Trainer  # noqa: E402,F401
