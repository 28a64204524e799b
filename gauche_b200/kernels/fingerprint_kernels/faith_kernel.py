"""Drop-in module path for gauche/kernels/fingerprint_kernels/faith_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import FaithKernel, batch_faith_sim

__all__ = ["FaithKernel", "batch_faith_sim"]
