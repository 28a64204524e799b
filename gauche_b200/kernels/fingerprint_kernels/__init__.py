"""Drop-in namespace mirroring ``gauche.kernels.fingerprint_kernels``
(gauche/kernels/fingerprint_kernels/__init__.py:1-31): the two hot-path kernels plus the twelve siblings."""
from .minmax_kernel import MinMaxKernel, batch_minmax_sim
from .sibling_kernels import (BraunBlanquetKernel, DiceKernel, FaithKernel, ForbesKernel, InnerProductKernel,
                              IntersectionKernel, OtsukaKernel, RandKernel, RogersTanimotoKernel,
                              RussellRaoKernel, SogenfreiKernel, SokalSneathKernel)
from .tanimoto_kernel import TanimotoKernel, batch_tanimoto_sim

__all__ = [
    "BraunBlanquetKernel", "DiceKernel", "FaithKernel", "ForbesKernel", "InnerProductKernel",
    "IntersectionKernel", "MinMaxKernel", "OtsukaKernel", "RandKernel", "RogersTanimotoKernel",
    "RussellRaoKernel", "SogenfreiKernel", "SokalSneathKernel", "TanimotoKernel",
    "batch_minmax_sim", "batch_tanimoto_sim",
]
