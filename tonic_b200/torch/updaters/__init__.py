from .actors import ClippedRatio, StochasticPolicyGradient
from .critics import VRegression

__all__ = [ClippedRatio, StochasticPolicyGradient, VRegression]
