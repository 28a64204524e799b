"""Drop-in module path for gauche/kernels/fingerprint_kernels/braun_blanquet_kernel.py (see sibling_kernels.py)."""
from .sibling_kernels import BraunBlanquetKernel, batch_braun_blanquet_sim

__all__ = ["BraunBlanquetKernel", "batch_braun_blanquet_sim"]
