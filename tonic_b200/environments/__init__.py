from .builders import SynthControl, Space
from .distributed import DeviceVectorEnvironment, distribute
from .host import HostParallel, HostSequential, distribute_host

__all__ = [SynthControl, Space, DeviceVectorEnvironment, distribute, HostSequential, HostParallel,
           distribute_host]
