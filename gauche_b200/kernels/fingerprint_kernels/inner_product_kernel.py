"""Drop-in module path for gauche/kernels/fingerprint_kernels/inner_product_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import InnerProductKernel, batch_inner_product_sim

__all__ = ["InnerProductKernel", "batch_inner_product_sim"]
