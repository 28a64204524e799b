"""Drop-in namespace mirroring ``gauche.kernels.fingerprint_kernels`` for the two hot-path kernels
(gauche/kernels/fingerprint_kernels/__init__.py:1-31)."""
from .minmax_kernel import MinMaxKernel, batch_minmax_sim
from .tanimoto_kernel import TanimotoKernel, batch_tanimoto_sim

__all__ = ["MinMaxKernel", "TanimotoKernel", "batch_minmax_sim", "batch_tanimoto_sim"]
