"""Dense fused gradient entry points (csrc/kernels/dense_round.cu)."""
from .grad import grad, grad_full  # noqa: F401
