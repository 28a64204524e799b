"""Drop-in module path for gauche/kernels/fingerprint_kernels/forbes_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import ForbesKernel, batch_forbes_sim

__all__ = ["ForbesKernel", "batch_forbes_sim"]
