"""Drop-in module path for gauche/kernels/fingerprint_kernels/intersection_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import IntersectionKernel, batch_intersection_sim

__all__ = ["IntersectionKernel", "batch_intersection_sim"]
