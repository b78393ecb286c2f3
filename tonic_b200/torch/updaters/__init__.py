from .actors import (ClippedRatio, DeterministicPolicyGradient, StochasticPolicyGradient,
                     TwinCriticSoftDeterministicPolicyGradient)
from .critics import (DeterministicQLearning, TargetActionNoise, TwinCriticDeterministicQLearning,
                      TwinCriticSoftQLearning, VRegression)

__all__ = [
    ClippedRatio, DeterministicPolicyGradient, StochasticPolicyGradient,
    TwinCriticSoftDeterministicPolicyGradient, DeterministicQLearning, TargetActionNoise,
    TwinCriticDeterministicQLearning, TwinCriticSoftQLearning, VRegression]
