"""CSR fused gradient entry points (csrc/kernels/csr_round.cu)."""
from .grad import grad, grad_full  # noqa: F401
