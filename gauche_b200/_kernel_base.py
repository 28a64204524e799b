"""Base class of the drop-in kernels.

With GPyTorch installed the kernels subclass the real ``gpytorch.kernels.Kernel`` and are used
under ExactGP / ScaleKernel / InducingPointKernel / BoTorch unchanged.  GPyTorch is not part of
this image, so a protocol-faithful stand-in is provided: it mirrors what ``Kernel.__call__``
does before reaching ``forward`` (1-D unsqueeze, ``x2 = x1`` default, ``diag`` short-cut, lazy
evaluation with ``.size()`` computed from the input shapes without running the kernel), which is
exactly what the reference's own shape tests exercise
(tests/test_kernels/test_fingerprint_kernels/test_tanimoto_kernel.py:82-96).
"""
from __future__ import annotations

import torch

try:  # pragma: no cover - depends on the environment
    from gpytorch.kernels import Kernel as Kernel  # type: ignore
    from gpytorch.kernels import ScaleKernel as ScaleKernel  # type: ignore

    HAVE_GPYTORCH = True
except Exception:  # gpytorch absent: protocol stand-in
    HAVE_GPYTORCH = False

    class LazyEvaluatedKernelTensor:
        """Deferred ``kernel(x1, x2)``: shape is known up front, values on ``to_dense()``."""

        def __init__(self, x1, x2, kernel, last_dim_is_batch=False, **params):
            self.x1, self.x2, self.kernel = x1, x2, kernel
            self.last_dim_is_batch = last_dim_is_batch
            self.params = params

        def size(self, dim=None):
            b = torch.broadcast_shapes(self.x1.shape[:-2], self.x2.shape[:-2])
            b = torch.broadcast_shapes(b, tuple(getattr(self.kernel, "batch_shape", torch.Size())))
            if self.last_dim_is_batch:
                b = tuple(b) + (self.x1.shape[-1],)
            full = torch.Size(tuple(b) + (self.x1.shape[-2], self.x2.shape[-2]))
            return full if dim is None else full[dim]

        @property
        def shape(self):
            return self.size()

        def to_dense(self):
            return self.kernel.forward(self.x1, self.x2, diag=False,
                                       last_dim_is_batch=self.last_dim_is_batch, **self.params)

        evaluate = to_dense
        evaluate_kernel = to_dense

        def diagonal(self):
            return self.kernel.forward(self.x1, self.x2, diag=True,
                                       last_dim_is_batch=self.last_dim_is_batch, **self.params)

        def __getitem__(self, index):
            # row / column slicing re-invokes the kernel on sliced inputs (cross-kernel protocol,
            # gauche/gp.py:197-226)
            if not isinstance(index, tuple):
                index = (index,)
            if len(index) != 2 or not all(isinstance(i, slice) for i in index):
                return self.to_dense()[index]
            return LazyEvaluatedKernelTensor(self.x1[..., index[0], :], self.x2[..., index[1], :],
                                             self.kernel, self.last_dim_is_batch, **self.params)

    class Kernel(torch.nn.Module):
        has_lengthscale = False
        is_stationary = False

        def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None,
                     lengthscale_prior=None, lengthscale_constraint=None, eps=1e-6, **kwargs):
            super().__init__()
            self._batch_shape = torch.Size(batch_shape)
            if active_dims is not None and not torch.is_tensor(active_dims):
                active_dims = torch.tensor(active_dims, dtype=torch.long)
            self.active_dims = active_dims
            self.ard_num_dims = ard_num_dims
            self.eps = eps

        @property
        def batch_shape(self):
            return self._batch_shape

        def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **params):
            raise NotImplementedError

        def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
            x1_, x2_ = x1, x2
            if self.active_dims is not None:
                x1_ = x1_.index_select(-1, self.active_dims.to(x1_.device))
                if x2_ is not None:
                    x2_ = x2_.index_select(-1, self.active_dims.to(x2_.device))
            if x1_.ndimension() == 1:
                x1_ = x1_.unsqueeze(1)
            if x2_ is not None:
                if x2_.ndimension() == 1:
                    x2_ = x2_.unsqueeze(1)
                if not x1_.size(-1) == x2_.size(-1):
                    raise RuntimeError("x1_ and x2_ must have the same number of dimensions!")
            if x2_ is None:
                x2_ = x1_
            if diag:
                return self.forward(x1_, x2_, diag=True, last_dim_is_batch=last_dim_is_batch, **params)
            return LazyEvaluatedKernelTensor(x1_, x2_, kernel=self, last_dim_is_batch=last_dim_is_batch,
                                             **params)

    class ScaleKernel(Kernel):
        """outputscale * base_kernel, calling ``base_kernel.forward`` directly like GPyTorch's."""

        def __init__(self, base_kernel, **kwargs):
            super().__init__(**kwargs)
            self.base_kernel = base_kernel
            self.raw_outputscale = torch.nn.Parameter(torch.zeros(tuple(self.batch_shape)))

        @property
        def outputscale(self):
            return torch.nn.functional.softplus(self.raw_outputscale)

        def forward(self, x1, x2, last_dim_is_batch=False, diag=False, **params):
            orig = self.base_kernel.forward(x1, x2, diag=diag, last_dim_is_batch=last_dim_is_batch, **params)
            scale = self.outputscale
            if last_dim_is_batch:
                scale = scale.unsqueeze(-1)
            if diag:
                return orig * scale.unsqueeze(-1).to(orig.dtype)
            return orig * scale.view(*scale.shape, 1, 1).to(orig.dtype)
