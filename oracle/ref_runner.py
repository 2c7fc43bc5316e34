"""Time the UNMODIFIED reference sampler / predict on host cores (bench.py's CPU legs only).

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.  The reference modules come from
`/root/reference/src` (build container) or the byte-identical staged copy under `baseline/_ref/src`
(`oracle/build_ref.py`), through `oracle/ref_shim.py`.

Two ways of driving the reference, both executing ITS njit code:
  * `RefChain.full_sweep()`  -- one call of `_step_bark_sampler` (src/bark/fitting/bark_sampler.py:216-284): m tree
    proposals + the noise/scale proposal, exactly as `_run_bark_sampler_multichain` (:120-213) does per sweep.
  * `RefChain.advance(t0, t1, hyper)` -- the same statements, in the same order, issued from Python one proposal at a
    time (`get_tree_proposal`, `get_leaf_vectors` x2, `low_rank_inv_update` x2, `low_rank_det_update` x2, `mll`, MH
    test; then `get_noise_scale_proposal`, `forest_gram_matrix`, `inv`, `slogdet`, `mll`), so that a bench step can be a
    BOUNDED slice of a sweep (a whole sweep of one chain is ~40 s at N=2000, m=200).  Python overhead per proposal
    is microseconds against ~0.2 s of numba/LAPACK work.
"""
from __future__ import annotations

import os
import time
import warnings

import numpy as np

from . import ref_shim


def available() -> bool:
    return ref_shim.available()


_mods = None


def _load():
    global _mods
    if _mods is None:
        warnings.filterwarnings("ignore")
        ref_shim.install()
        import bark.forest as RF
        from bark.fitting import bark_sampler as RS
        from bark.fitting import noise_scale_proposals as RN
        from bark.fitting import quick_inverse as RQ
        from bark.fitting import tree_proposals as RT
        _mods = dict(RF=RF, RS=RS, RN=RN, RQ=RQ, RT=RT)
    return _mods


def ref_params(num_chains=1, warmup_steps=0, num_samples=1, steps_per_sample=1, alpha=0.95, beta=2.0,
               proposal_weights=(0.25, 0.25, 0.5), use_softplus_transform=True, sample_scale=False,
               gamma_prior_shape=1.5, gamma_prior_rate=5.0):
    RS = _load()["RS"]
    return RS.BARKTrainParamsNumba(warmup_steps, num_samples, steps_per_sample, num_chains, alpha, beta,
                                   np.asarray(proposal_weights, dtype=np.float64), False, use_softplus_transform,
                                   sample_scale, gamma_prior_shape, gamma_prior_rate)


def set_blas_threads(n: int | None):
    """Thread count of every BLAS / OpenMP pool in the process (torchrun exports OMP_NUM_THREADS=1, which would
    silently halve the reference's LAPACK-bound rate).  Returns the threadpoolctl limiter (keep it alive)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n or (os.cpu_count() or 1))
    except Exception:  # threadpoolctl missing: the environment decides
        return None


class RefChain:
    """One chain of the reference sampler, state as in `_run_bark_sampler_multichain` (:147-162)."""

    def __init__(self, forest, noise, scale, X, y, bounds, feat_types, params=None):
        M = _load()
        self.M = M
        self.forest = np.ascontiguousarray(forest).copy()
        self.noise, self.scale = float(noise), float(scale)
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1, 1)
        self.bounds = np.ascontiguousarray(bounds, dtype=np.float64)
        self.ft = np.ascontiguousarray(feat_types, dtype=np.int64)
        self.params = params if params is not None else ref_params()
        K = self.scale * M["RF"].forest_gram_matrix(self.forest, self.X, self.X, self.ft)
        K = K + (1e-6 + self.noise) * np.eye(K.shape[0])
        self.K_inv = np.linalg.inv(K)
        self.logdet = np.linalg.slogdet(K)[1]
        self.mll = M["RQ"].mll(self.K_inv, self.logdet, self.y)

    def full_sweep(self) -> int:
        RS = self.M["RS"]
        (self.forest, self.noise, self.scale, self.K_inv, self.logdet, self.mll) = RS._step_bark_sampler(
            self.forest, self.noise, self.scale, self.X, self.y, self.bounds, self.ft, self.params, self.K_inv,
            self.logdet, self.mll)
        return self.forest.shape[0] + 1

    def advance(self, t_begin: int, t_end: int, do_hyper: bool) -> int:
        """Tree proposals [t_begin, t_end) of a sweep, then optionally the noise/scale proposal; returns the number
        of proposals evaluated (each with its full log-MLL)."""
        M = self.M
        RF, RQ, RT, RN = M["RF"], M["RQ"], M["RT"], M["RN"]
        m = self.forest.shape[0]
        s_sqrtm = np.sqrt(self.scale / m)
        done = 0
        for t in range(t_begin, t_end):
            new_nodes, lqp = RT.get_tree_proposal(self.forest[t], self.bounds, self.ft, self.params)
            cur_lv = s_sqrtm * RF.get_leaf_vectors(self.forest[t], self.X, self.ft)
            new_lv = s_sqrtm * RF.get_leaf_vectors(new_nodes, self.X, self.ft)
            k_inv = RQ.low_rank_inv_update(self.K_inv, cur_lv, subtract=True)
            ld = RQ.low_rank_det_update(self.K_inv, cur_lv, self.logdet, subtract=True)
            k_inv2 = RQ.low_rank_inv_update(k_inv, new_lv)
            ld2 = RQ.low_rank_det_update(k_inv, new_lv, ld)
            new_mll = RQ.mll(k_inv2, ld2, self.y)
            if np.log(np.random.uniform()) <= min(lqp + new_mll - self.mll, 0):
                self.K_inv, self.logdet, self.mll = k_inv2, ld2, new_mll
                self.forest[t] = new_nodes
            done += 1
        if do_hyper:
            (nn, ns), lqp = RN.get_noise_scale_proposal(self.noise, self.scale, self.params)
            K = ns * RF.forest_gram_matrix(self.forest, self.X, self.X, self.ft)
            K = K + (1e-6 + nn) * np.eye(K.shape[0])
            k_inv = np.linalg.inv(K)
            ld = np.linalg.slogdet(K)[1]
            new_mll = RQ.mll(k_inv, ld, self.y)
            if np.log(np.random.uniform()) <= min(lqp + new_mll - self.mll, 0):
                self.K_inv, self.logdet, self.mll, self.noise, self.scale = k_inv, ld, new_mll, nn, ns
            done += 1
        return done


def warm_jit(d: int, cat: bool = False):
    """Compile every njit function on a tiny problem (excluded from all timings; ~1 min the first time)."""
    from . import bark_oracle as O
    Xs, ys, bs, fs, _ = O.synthetic_problem(24, dim=d if not cat else max(1, d - 1), cat_dim=1 if cat else 0,
                                            num_cat=3, m_true=4, seed=0)
    ch = RefChain(O.create_empty_forest(3), 0.1, 1.0, Xs, ys, bs, fs)
    ch.advance(0, 3, True)
    ch.full_sweep()


def time_sampler_slices(forest, noise, scale, X, y, bounds, ft, budget_s: float, steps: int = 1, warmup: int = 0,
                        threads: int | None = None):
    """Bounded sample: (proposals/s, per-step seconds, proposals per step, description).  Each step is a slice of one
    chain's sweep sized so that warmup + steps fit `budget_s`."""
    lim = set_blas_threads(threads)
    ch = RefChain(forest, noise, scale, X, y, bounds, ft)  # K^-1 build: untimed, like chain init on the GPU
    m = forest.shape[0]
    k0 = min(3, m)
    t0 = time.perf_counter(); ch.advance(0, k0, False); probe = (time.perf_counter() - t0) / k0
    per_step = max(1, min(m, int(budget_s / max(probe, 1e-9) / max(steps + warmup, 1))))
    times, props, cur = [], 0, k0 % m
    for it in range(warmup + steps):
        t_end = min(m, cur + per_step)
        hyper = t_end == m
        t0 = time.perf_counter()
        k = ch.advance(cur, t_end, hyper)
        dt = time.perf_counter() - t0
        cur = 0 if t_end == m else t_end
        if it >= warmup:
            times.append(dt); props += k
    del lim
    return props / sum(times), times, per_step, props


def time_full_sweep(forest, noise, scale, X, y, bounds, ft, threads: int | None = None):
    """One whole `_step_bark_sampler` call of one chain: (proposals/s, seconds)."""
    lim = set_blas_threads(threads)
    ch = RefChain(forest, noise, scale, X, y, bounds, ft)
    t0 = time.perf_counter(); k = ch.full_sweep(); dt = time.perf_counter() - t0
    del lim
    return k / dt, dt


def time_predict_chunk(model, data, candidates, feat_types, threads: int | None = None):
    """The reference's `forest_predict` (src/bark/tree_kernels/tree_gps.py:80-113) on a candidate chunk:
    (points/s for the mixture over all samples, seconds).  The n_c x n_c covariance it forms bounds the chunk."""
    _load()
    from bark.tree_kernels import tree_gps as TG
    lim = set_blas_threads(threads)
    ft = np.ascontiguousarray(feat_types, dtype=np.int64)
    TG.get_feature_types_array = lambda domain: ft  # the only bofire-dependent call on this path (:94)
    t0 = time.perf_counter()
    mu, var = TG.forest_predict(model, data, np.ascontiguousarray(candidates, dtype=np.float64), None, diag=True)
    TG.mixture_of_gaussians_as_normal(mu, var)
    dt = time.perf_counter() - t0
    del lim
    return candidates.shape[0] / dt, dt
