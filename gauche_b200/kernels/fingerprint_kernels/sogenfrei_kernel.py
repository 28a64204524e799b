"""Drop-in module path for gauche/kernels/fingerprint_kernels/sogenfrei_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import SogenfreiKernel, batch_sogenfrei_sim

__all__ = ["SogenfreiKernel", "batch_sogenfrei_sim"]
