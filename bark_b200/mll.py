"""Batched FP64 log-marginal-likelihood on the GPU (a5).

mll = 0.5 * (-y^T K^-1 y - log|K|)   (src/bark/fitting/quick_inverse.py:36-38), evaluated from scratch for a
batch of kernel matrices as at src/bark/fitting/bark_sampler.py:153-162."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .forest import (DeviceForest, _as_device_f64, _feat_types_device, _ptr, _stream, forest_slots, gram_umma_device,
                     traverse_device)


def mll_batched_device(K, y):
    """K (B, N, N) f64 CUDA tensor (destroyed), y (N,) f64 CUDA tensor -> (mll, logdet, quad, status) tensors."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    b, n, _ = K.shape
    out = torch.empty((3, b), dtype=torch.float64, device=K.device)
    status = torch.zeros(b, dtype=torch.int32, device=K.device)
    ws = torch.empty(max(int(lib.bark_mll_workspace_bytes(b, n)), 8), dtype=torch.uint8, device=K.device)
    _lib.check(lib.bark_mll_batched(_ptr(K), b, n, _ptr(y), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(status),
                                    _ptr(ws), _stream()))
    return out[0], out[1], out[2], status


def forest_mll(nodes: np.ndarray, noise, scale, X: np.ndarray, y: np.ndarray, feat_types) -> np.ndarray:
    """Full log-MLL of a batch of forests (B, m, L) with hyper-parameters noise/scale (B,):
    traverse -> int8 tcgen05 Gram with fused FP64 epilogue K = scale*K0 + (1e-6+noise) I -> batched block LDL^T."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda")
    nodes = nodes.reshape(-1, *nodes.shape[-2:])
    df = DeviceForest.from_numpy(nodes, dev)
    Xd = _as_device_f64(X, dev)
    leaves = traverse_device(df, Xd, _feat_types_device(feat_types, dev))
    _, K = gram_umma_device(leaves, leaves, slots=forest_slots(nodes), want_counts=False,
                            scale=_as_device_f64(np.reshape(scale, -1), dev), noise=_as_device_f64(np.reshape(noise, -1), dev))
    val, _, _, status = mll_batched_device(K, _as_device_f64(np.reshape(y, -1), dev))
    if int(status.max().item()) & _lib.ST_NOT_SPD:
        raise np.linalg.LinAlgError("kernel matrix is not positive definite")
    return val.cpu().numpy()
