"""Drop-in module path for gauche/kernels/fingerprint_kernels/rogers_tanimoto_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import RogersTanimotoKernel, batch_rogers_tanimoto_sim

__all__ = ["RogersTanimotoKernel", "batch_rogers_tanimoto_sim"]
