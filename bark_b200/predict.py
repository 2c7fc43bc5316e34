"""Posterior predictive: `forest_predict` / `mixture_of_gaussians_as_normal`
(src/bark/tree_kernels/tree_gps.py:80-131) on the GPU, in leaf space (csrc/predict.cu)."""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import numpy as np

from . import _lib
from .domain import unpack_domain
from .forest import _as_device_f64, _ptr, _stream, forest_slots
from .sampler import ChainState, raise_for_status


class BARKModel(NamedTuple):  # src/bark/tree_kernels/tree_gps.py:14-17
    forest: np.ndarray
    noise: np.ndarray
    scale: np.ndarray


class PosteriorState:
    """Per-sample leaf-space state (B^-1, w = B^-1 b, column maps) of all posterior samples, resident in HBM.
    Build once, predict many candidate batches."""


    def __init__(self, model, data, feat_types, d, p_cap=None, device=None, tensor_cores=True):
        forest, noise, scale = model
        forest = np.ascontiguousarray(forest).reshape(-1, *forest.shape[-2:])
        noise = np.asarray(noise, dtype=np.float64).reshape(-1)
        scale = np.asarray(scale, dtype=np.float64).reshape(-1)
        train_x, train_y = data
        if p_cap is None:
            leaves = int((forest["active"] & forest["is_leaf"]).sum(axis=(-1, -2)).max())
            p_cap = max(64, ((leaves + 63) // 64) * 64)
        bounds = np.zeros((d, 2))
        self.state = ChainState(forest, noise, scale, train_x, train_y, bounds, feat_types, p_cap=p_cap, device=device)
        self.num_samples = forest.shape[0]
        # tensor-core path: slice every sample's B^-1 into int8 digit planes once (leaf-column extent <= 768)
        st = self.state
        self.slots = forest_slots(forest)
        self.p_max = int(st.read()["p_used"].max().item())
        self.prep = None
        self.k_pad = None  # one-hot width of the int8 GEMM (multiple of 128); see umma_ops_per_candidate_sample
        # (extents above 768 columns, or tensor_cores=False, use the FP64 gather kernel of csrc/predict.cu)
        nbytes = int(st.lib.bark_predict_prep_bytes(C.byref(st.dims), self.slots, self.p_max)) if (tensor_cores and self.p_max <= 768) else 0
        if nbytes:
            torch = _lib.require_cuda()
            self.prep = torch.empty(nbytes, dtype=torch.uint8, device=st.device)
            self.k_pad = ((self.p_max + 127) // 128) * 128
            _lib.check(st.lib.bark_predict_prepare(C.byref(st.dims), _ptr(st.ws), st.dforest.soa(), self.slots, self.p_max,
                                                   _ptr(self.prep), _stream()))

    @property
    def umma_ops_per_candidate_sample(self):
        """int8 multiply-adds x 2 that the tensor-core path executes per (candidate, posterior sample): 7 digit planes of
        the triangular operand -- column tiles of 192 (160 above 512 columns; the last may be narrower), K tiles of 128 from the
        diagonal band down
        (csrc/predict_umma.cu, pu_kt_lo)."""
        if not self.k_pad:
            return None
        kt = self.k_pad // 128
        ntile = 192 if kt <= 4 else 160  # pu_ntile: accumulator width beside the A operand in TMEM
        ops, nt = 0, 0
        while nt * ntile < self.k_pad:
            ncols = min(ntile, self.k_pad - nt * ntile)
            ops += 2 * 7 * ncols * 128 * (kt - (nt * ntile) // 128)
            nt += 1
        return ops

    def check(self):
        raise_for_status(self.state.read()["status"].cpu().numpy())

    def predict_cov_device(self, cand_dev):
        """cand_dev (n_c, d) -> (mu (S, n_c), cov (S, n_c, n_c)): `forest_predict(..., diag=False)`."""
        torch = _lib.require_cuda()
        st = self.state
        n_c = cand_dev.shape[0]
        mu = torch.empty((self.num_samples, n_c), dtype=torch.float64, device=st.device)
        cov = torch.empty((self.num_samples, n_c, n_c), dtype=torch.float64, device=st.device)
        scratch = torch.empty(max(int(st.lib.bark_predict_cov_scratch_bytes(C.byref(st.dims), n_c)), 8), dtype=torch.uint8,
                              device=st.device)
        _lib.check(st.lib.bark_predict_cov(C.byref(st.dims), _ptr(st.ws), st.dforest.soa(), _ptr(cand_dev), n_c, _ptr(mu),
                                           _ptr(cov), _ptr(scratch), _stream()))
        return mu, cov

    def predict_device(self, cand_dev, mode=0, y_mean=0.0, y_std=1.0, add_noise=False):
        """cand_dev (n_c, d) f64 CUDA tensor -> (mu, var) CUDA tensors: (S, n_c) for mode 0, (n_c,) for mode 1."""
        torch = _lib.require_cuda()
        st = self.state
        n_c = cand_dev.shape[0]
        shape = (self.num_samples, n_c) if mode == 0 else (n_c,)
        mu = torch.empty(shape, dtype=torch.float64, device=st.device)
        var = torch.empty(shape, dtype=torch.float64, device=st.device)
        if self.prep is not None:  # int8 tcgen05 path
            if mode == 0:
                _lib.check(st.lib.bark_predict_umma(C.byref(st.dims), _ptr(st.ws), _ptr(self.prep), self.slots, self.p_max,
                                                    _ptr(cand_dev), n_c, _ptr(mu), _ptr(var), _stream()))
                return mu, var
            # mixture mode: the persistent kernel folds the samples in registers (no per-sample moments in memory)
            _lib.check(st.lib.bark_predict_umma_mixture(C.byref(st.dims), _ptr(st.ws), _ptr(self.prep), self.slots, self.p_max,
                                                        _ptr(cand_dev), n_c, float(y_mean), float(y_std),
                                                        int(bool(add_noise)), _ptr(mu), _ptr(var), _stream()))
            return mu, var
        nbytes = int(st.lib.bark_predict_scratch_bytes(C.byref(st.dims), n_c))
        scratch = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=st.device)
        _lib.check(st.lib.bark_predict(C.byref(st.dims), _ptr(st.ws), st.dforest.soa(), _ptr(cand_dev), n_c, int(mode),
                                       float(y_mean), float(y_std), int(bool(add_noise)), _ptr(mu), _ptr(var),
                                       _ptr(scratch), _stream()))
        return mu, var


def forest_predict(model, data, candidates: np.ndarray, domain, diag: bool = True):
    """Per-sample posterior mean and variance at `candidates` (src/bark/tree_kernels/tree_gps.py:80-113).
    Leading sample dims of the model are flattened; returns (mu (S_tot, n_c), var (S_tot, n_c))."""
    _lib.require_cuda()
    _, feat_types = unpack_domain(domain)
    candidates = np.ascontiguousarray(candidates, dtype=np.float64)
    if not diag:  # the full (S_tot, n_c, n_c) matrix, as the reference forms it (tree_gps.py:107-112)
        ps = PosteriorState(model, data, feat_types, candidates.shape[1], tensor_cores=False)
        mu, cov = ps.predict_cov_device(_as_device_f64(candidates, ps.state.device))
        ps.check()
        return mu.cpu().numpy(), cov.cpu().numpy()
    ps = PosteriorState(model, data, feat_types, candidates.shape[1])
    mu, var = ps.predict_device(_as_device_f64(candidates, ps.state.device), mode=0)
    ps.check()
    return mu.cpu().numpy(), var.cpu().numpy()


def mixture_of_gaussians_as_normal(mu: np.ndarray, var: np.ndarray):
    """Mean and variance of an equal-weight mixture of Gaussians (src/bark/tree_kernels/tree_gps.py:116-131).
    (Host version for arrays already on the host; `PosteriorState.predict_device(mode=1)` fuses it on the GPU.)"""
    mu_y = np.mean(mu, axis=0)
    var_y = np.mean(var + mu**2, axis=0) - mu_y**2
    return mu_y, var_y
