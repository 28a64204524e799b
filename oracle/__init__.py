"""CPU oracle of the fingerprint-kernel Gram path — TEST INFRASTRUCTURE ONLY.

Nothing under ``gauche_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may use it, and only as the checker / the timed CPU baseline.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference functions
(loaded from /root/reference with a stub ``gpytorch`` module) in the build container and commits
their inputs/outputs to ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
oracle function bit-for-bit against those vectors and against the reference's own known-answer
tests (tests/test_kernels/test_fingerprint_kernels/test_tanimoto_kernel.py:18-79,
test_minmax_kernel.py:17-38).
"""
from .exact import (minmax_l1, minmax_sim, minmax_sim_from_counts, tanimoto_counts, tanimoto_sim,
                    tanimoto_sim_from_counts)
from .torch_port import minmax_sim_torch, tanimoto_sim_torch

__all__ = [
    "tanimoto_counts", "tanimoto_sim_from_counts", "tanimoto_sim", "minmax_l1",
    "minmax_sim_from_counts", "minmax_sim", "tanimoto_sim_torch", "minmax_sim_torch",
]
