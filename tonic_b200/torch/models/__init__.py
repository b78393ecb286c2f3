from .actor_critics import ActorCritic, ActorCriticWithTargets, ActorTwinCriticWithTargets
from .actors import (Actor, DetachedScaleGaussianPolicyHead, DeterministicPolicyHead,
                     GaussianPolicyHead, SquashedMultivariateNormalDiag)
from .critics import Critic, ValueHead
from .encoders import ObservationActionEncoder, ObservationEncoder
from .utils import MLP, trainable_variables

__all__ = [
    MLP, trainable_variables, ObservationActionEncoder, ObservationEncoder,
    SquashedMultivariateNormalDiag, DetachedScaleGaussianPolicyHead, GaussianPolicyHead,
    DeterministicPolicyHead, Actor, Critic, ValueHead, ActorCritic, ActorCriticWithTargets,
    ActorTwinCriticWithTargets]
