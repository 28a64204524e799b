"""Drop-in module path for gauche/kernels/fingerprint_kernels/russell_rao_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import RussellRaoKernel, batch_russell_rao_sim

__all__ = ["RussellRaoKernel", "batch_russell_rao_sim"]
