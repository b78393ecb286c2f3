from .noisy import NoActionNoise, NormalActionNoise

__all__ = [NoActionNoise, NormalActionNoise]
