"""Drop-in module path for gauche/kernels/fingerprint_kernels/otsuka_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import OtsukaKernel, batch_otsuka_sim

__all__ = ["OtsukaKernel", "batch_otsuka_sim"]
