from .builders import SynthControl, Space
from .classic import Gym
from .distributed import DeviceVectorEnvironment, distribute
from .host import HostParallel, HostSequential, distribute_host

__all__ = [SynthControl, Space, Gym, DeviceVectorEnvironment, distribute, HostSequential, HostParallel,
           distribute_host]
