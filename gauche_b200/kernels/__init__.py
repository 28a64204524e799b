from .fingerprint_kernels import MinMaxKernel, TanimotoKernel

__all__ = ["MinMaxKernel", "TanimotoKernel"]
