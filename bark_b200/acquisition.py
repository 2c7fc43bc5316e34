"""Inputs of the acquisition model: what `build_opt_model_from_forest`
(src/bark/optimizer/opt_model.py:31-110) computes on the host with a batched `np.linalg.inv` right after the fit --
per posterior sample `K_inv`, the quadratic term `-scale^2 K_inv` and the linear term `scale K_inv y` -- served
from the GPU's leaf-space state by Woodbury (csrc/kinv.cu).  The MIP itself (gurobipy) stays on the host and is
out of scope."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .domain import unpack_domain
from .forest import _as_device_f64, _feat_types_device, _ptr, _stream, traverse_device
from .sampler import ChainState, raise_for_status


def gp_sample_inverses(gp_samples, data, domain, standardize_y=True, device=False, p_cap=None):
    """gp_samples = (forest (..., m, L), noise (...), scale (...)), data = (X (N, D), y (N, 1)).

    Returns dict(K_inv (S, N, N), quadr_term (S, N, N) = -scale^2 K_inv, lin_term (S, N) = scale K_inv y,
    const_term (S,) = scale) with the sample dims flattened (opt_model.py:36-43); numpy arrays, or CUDA tensors
    with device=True.  The kernel is the one of `batched_forest_gram_matrix_no_null` (root-only trees removed),
    and y is re-standardised as at opt_model.py:28."""
    torch = _lib.require_cuda()
    forest, noise, scale = gp_samples
    forest = np.ascontiguousarray(forest).reshape(-1, *forest.shape[-2:])
    noise = np.asarray(noise, dtype=np.float64).reshape(-1)
    scale = np.asarray(scale, dtype=np.float64).reshape(-1)
    X, y = data
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    if standardize_y:
        y = (y - y.mean()) / y.std()
    _, feat_types = unpack_domain(domain)
    S, m = forest.shape[0], forest.shape[1]
    n = X.shape[0]
    # the no-null kernel is (scale / m') * Z' Z'^T + sig I with m' = number of trees that split at least once
    if p_cap is None:
        leaves_max = int((forest["active"] & forest["is_leaf"]).sum(axis=(-1, -2)).max())
        p_cap = max(64, ((leaves_max + 63) // 64) * 64)
    st = ChainState(forest, noise, scale, X, y, np.zeros((X.shape[1], 2)), feat_types, p_cap=p_cap, skip_null=True)
    raise_for_status(st.read()["status"].cpu().numpy())
    leaves = traverse_device(st.dforest, st.X, st.ft)  # (S, n, m) uint32 slot ids
    dev = st.device
    kinv = torch.empty((S, n, n), dtype=torch.float64, device=dev)
    kinv_y = torch.empty((S, n), dtype=torch.float64, device=dev)
    scratch = torch.empty(max(int(st.lib.bark_kinv_scratch_bytes(C.byref(st.dims))), 8), dtype=torch.uint8, device=dev)
    _lib.check(st.lib.bark_kinv_export(C.byref(st.dims), _ptr(st.ws), _ptr(leaves), _ptr(kinv), _ptr(kinv_y),
                                       _ptr(scratch), _stream()))
    sc = _as_device_f64(scale, dev)
    out = {"K_inv": kinv, "quadr_term": -(sc * sc)[:, None, None] * kinv, "lin_term": sc[:, None] * kinv_y,
           "const_term": sc}
    if device:
        return out
    return {k: v.cpu().numpy() for k, v in out.items()}
