"""gauche_b200 — B200-native (sm_100a) Tanimoto / MinMax fingerprint-kernel Gram matrices.

Drop-in for the hot path of leojklarner/gauche:
``gauche.kernels.fingerprint_kernels.{TanimotoKernel, MinMaxKernel, batch_tanimoto_sim, batch_minmax_sim}``.
"""
from .kernels.fingerprint_kernels import (MinMaxKernel, TanimotoKernel, batch_minmax_sim,
                                          batch_tanimoto_sim)
from .ops import intersection_counts, l1_distances, minmax_similarity, tanimoto_similarity
from .packed import PackedFingerprints, pack, pack_cache

__version__ = "0.1.0"
__all__ = [
    "MinMaxKernel", "TanimotoKernel", "batch_minmax_sim", "batch_tanimoto_sim",
    "tanimoto_similarity", "minmax_similarity", "intersection_counts", "l1_distances",
    "PackedFingerprints", "pack", "pack_cache",
]
