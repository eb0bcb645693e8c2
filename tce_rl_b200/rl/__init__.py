"""Host-side mirrors of ``mprl.rl`` for the policy-update path (policy / projection / critic / agent)."""
from .policy import BlackBoxPolicy, TemporalCorrelatedPolicy, policy_factory  # noqa: F401
from .projection import (BaseProjectionLayer, FrobeniusProjectionLayer, KLProjectionLayer,  # noqa: F401
                         WassersteinProjectionLayer, projection_factory)
from .critic import ValueFunction, critic_factory  # noqa: F401
from .agent import BlackBoxAgent, TemporalCorrelatedAgent, agent_factory  # noqa: F401
