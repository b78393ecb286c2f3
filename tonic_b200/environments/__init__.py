from .builders import SynthControl, Space
from .distributed import DeviceVectorEnvironment, distribute

__all__ = [SynthControl, Space, DeviceVectorEnvironment, distribute]
