from .derivative import derivative, num_derivative
from .phase import MomentumPhase
