"""Drop-in module path for gauche/kernels/fingerprint_kernels/sokal_sneath_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import SokalSneathKernel, batch_sokal_sneath_sim

__all__ = ["SokalSneathKernel", "batch_sokal_sneath_sim"]
