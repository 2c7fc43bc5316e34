"""`TreeAgreementKernel` (src/bark/tree_kernels/tree_model_kernel.py:8-23) served by the tcgen05 Gram: the fraction of
trees in which two points share a leaf, for tensors that already live on the GPU (no `.numpy()` round trip).
gpytorch is not required: the class is a plain callable with the reference's `forward(x1, x2, diag=False)`; when
gpytorch is importable it can be wrapped in a `gpytorch.kernels.Kernel` by the caller."""
from __future__ import annotations

import numpy as np

from . import _lib
from .forest import DeviceForest, _feat_types_device, forest_slots, gram_umma_device, traverse_device


class TreeAgreementKernel:
    is_stationary = False

    def __init__(self, forest: np.ndarray, feat_types: np.ndarray, device=None):
        torch = _lib.require_cuda()
        forest = np.ascontiguousarray(forest)
        if forest.ndim != 2:
            raise ValueError("forest must have shape (m, node_limit)")
        self.forest, self.feat_types = forest, np.asarray(feat_types)
        self.device = torch.device(device or "cuda")
        self._dforest = DeviceForest.from_numpy(forest[None], self.device)
        self._ft = _feat_types_device(self.feat_types, self.device)
        self._slots = forest_slots(forest[None])

    def forward(self, x1, x2, diag: bool = False, **params):
        """x1 (n1, d), x2 (n2, d) torch tensors (any device / float dtype) -> (n1, n2) float64 tensor on the GPU
        (or ones(n1) for diag=True, as in the reference)."""
        torch = _lib.require_cuda()
        if diag:
            return torch.ones(x1.shape[0], dtype=torch.float64, device=self.device)
        a = x1.detach().to(device=self.device, dtype=torch.float64).contiguous()
        b = a if x2 is x1 else x2.detach().to(device=self.device, dtype=torch.float64).contiguous()
        la = traverse_device(self._dforest, a, self._ft)
        lb = la if b is a else traverse_device(self._dforest, b, self._ft)
        ones = torch.ones(1, dtype=torch.float64, device=self.device)
        _, K0 = gram_umma_device(la, lb, slots=self._slots, want_counts=False, scale=ones)
        return K0[0]

    __call__ = forward
