"""Drop-in module path for gauche/kernels/fingerprint_kernels/rand_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import RandKernel, batch_rand_sim

__all__ = ["RandKernel", "batch_rand_sim"]
