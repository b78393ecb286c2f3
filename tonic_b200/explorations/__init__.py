from .noisy import NoActionNoise, NormalActionNoise, OrnsteinUhlenbeckActionNoise

__all__ = [NoActionNoise, NormalActionNoise, OrnsteinUhlenbeckActionNoise]
