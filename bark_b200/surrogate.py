"""bofire-free `BARKSurrogate`: same attributes, defaults and fit/predict behaviour as
`src/bofire_mixed/surrogates/bark.py:39-149` + `src/bofire_mixed/data_models/surrogates/bark.py:15-40`,
driving the GPU sampler / predictor."""
from __future__ import annotations

import numpy as np

from .domain import unpack_domain
from .forest import create_empty_forest
from .predict import PosteriorState
from .sampler import BARKTrainParams, run_bark_sampler
from .forest import _as_device_f64


class Standardize:  # src/bofire_mixed/standardize.py:4-21
    def __init__(self):
        self.mean = 0.0
        self.std = 1.0

    def __call__(self, y: np.ndarray, train: bool) -> np.ndarray:
        if train:
            self.mean = y.mean()
            self.std = max(y.std(), 1e-6)
        return (y - self.mean) / self.std

    def untransform(self, y):
        return y * self.std + self.mean

    def untransform_mu_var(self, mu, var):
        return self.untransform(mu), var * self.std**2


class BARKSurrogate:
    def __init__(self, domain, *, warmup_steps=50, num_samples=5, steps_per_sample=10, alpha=0.95, beta=2.0,
                 num_trees=50, use_softplus_transform=True, sample_scale=False, gamma_prior_shape=1.5,
                 gamma_prior_rate=5.0, grow_prune_weight=0.5, change_weight=1.0, num_chains=1, verbose=False, seed=None):
        self.domain = domain
        self.alpha, self.beta, self.num_trees = alpha, beta, num_trees
        self.gamma_prior_shape, self.gamma_prior_rate = gamma_prior_shape, gamma_prior_rate
        self.warmup_steps, self.num_samples, self.steps_per_sample = warmup_steps, num_samples, steps_per_sample
        self.num_chains, self.verbose = num_chains, verbose
        self.use_softplus_transform, self.sample_scale = use_softplus_transform, sample_scale
        w = np.array([grow_prune_weight, grow_prune_weight, change_weight], dtype=np.float64)
        self.bark_params = BARKTrainParams(  # _bark_params_to_jitclass, surrogates/bark.py:24-36
            warmup_steps=warmup_steps, num_samples=num_samples, steps_per_sample=steps_per_sample, num_chains=num_chains,
            alpha=alpha, beta=beta, proposal_weights=w / w.sum(), verbose=verbose,
            use_softplus_transform=use_softplus_transform, sample_scale=sample_scale,
            gamma_prior_shape=gamma_prior_shape, gamma_prior_rate=gamma_prior_rate)
        self.forest = None
        self.noise = None
        self.scale = None
        self.train_data = None
        self.scaler = Standardize()
        self._seed = seed
        self._fits = 0
        self._posterior = None

    def model_as_tuple(self):
        model = (self.forest, self.noise, self.scale)
        return None if any(x is None for x in model) else model

    @property
    def is_fitted(self) -> bool:
        return self.model_as_tuple() is not None

    def _init_bark(self):  # surrogates/bark.py:116-121
        forest = create_empty_forest(self.num_trees)
        self.forest = np.tile(forest, (self.num_chains, 1, 1, 1))
        self.noise = np.tile(0.1, (self.num_chains, 1))
        self.scale = np.tile(1.0, (self.num_chains, 1))

    def fit(self, X: np.ndarray, Y: np.ndarray):
        """X: (N, D) with categoricals ORDINAL-encoded; Y: (N, 1).  surrogates/bark.py:123-149."""
        Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
        self.train_data = (np.ascontiguousarray(X, dtype=np.float64), self.scaler(Y, train=True))
        if not self.is_fitted:
            self._init_bark()
        else:
            self.bark_params.warmup_steps = 0  # already warmed up: continue from the most recent sample
        most_recent = (np.ascontiguousarray(self.forest[:, -1, :, :]), self.noise[:, -1], self.scale[:, -1])
        seed = None if self._seed is None else self._seed + self._fits
        self.forest, self.noise, self.scale = run_bark_sampler(most_recent, self.train_data, self.domain,
                                                               self.bark_params, seed=seed)
        self._fits += 1
        self._posterior = None
        return self

    def predict(self, X: np.ndarray, batched=False, predict_observed=True):
        """(mu, std), each (n, 1) (or (S, n, 1) when batched).  surrogates/bark.py:71-94."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if self._posterior is None:
            _, ft = unpack_domain(self.domain)
            self._posterior = PosteriorState(self.model_as_tuple(), self.train_data, ft, X.shape[1])
            self._posterior.check()  # NOT_SPD / column overflow at state build must not yield silent garbage
        ps = self._posterior
        cand = _as_device_f64(X, ps.state.device)
        if batched:
            mu, var = ps.predict_device(cand, mode=0)
            mu, var = self.scaler.untransform_mu_var(mu.cpu().numpy(), var.cpu().numpy())
            if predict_observed:
                var = var + self.noise.reshape(-1, 1)
        else:
            mu, var = ps.predict_device(cand, mode=1, y_mean=self.scaler.mean, y_std=self.scaler.std,
                                        add_noise=predict_observed)
            mu, var = mu.cpu().numpy(), var.cpu().numpy()
        return mu[..., np.newaxis], np.sqrt(var[..., np.newaxis])


    # ---- on-disk samples (the reference's _dumps / loads are stubs, surrogates/bark.py:96-100)
    def save(self, path) -> None:
        from .checkpoint import save_samples
        save_samples(path, self.model_as_tuple(), self.train_data, self.bark_params,
                     extra={"scaler_mean": float(self.scaler.mean), "scaler_std": float(self.scaler.std), "fits": self._fits})

    def load(self, path):
        from .checkpoint import load_samples
        ck = load_samples(path)
        self.forest, self.noise, self.scale = ck["model"]
        self.train_data = ck["data"]
        self.scaler.mean, self.scaler.std = ck["extra"].get("scaler_mean", 0.0), ck["extra"].get("scaler_std", 1.0)
        self._fits = int(ck["extra"].get("fits", 1))
        self._posterior = None
        return self


class BARKPriorSurrogate(BARKSurrogate):
    """Samples from the BARK prior instead of the posterior (src/bofire_mixed/surrogates/bark.py:152-189):
    `fit` only stores the training data and draws `num_samples` prior forests / noise values; `predict` is the
    same GPU path as for posterior samples."""

    def __init__(self, domain, *, num_samples=5, sample_seed=0, gamma_prior_shape=2.5, gamma_prior_rate=9.0,
                 prior_on_device=False, **kwargs):
        # defaults of the reference data model (src/bofire_mixed/data_models/surrogates/bark.py:74-86):
        # inverse-gamma(2.5, 9.0) noise prior and a fixed sample_seed=0, not the posterior surrogate's 1.5 / 5.0
        super().__init__(domain, num_samples=num_samples, gamma_prior_shape=gamma_prior_shape,
                         gamma_prior_rate=gamma_prior_rate, **kwargs)
        self.sample_seed = sample_seed
        self.sample_rng = np.random.default_rng(sample_seed)
        self.prior_on_device = bool(prior_on_device)  # grow the forests with csrc/prior.cu instead of the host loop

    def fit(self, X: np.ndarray, Y: np.ndarray):
        from .prior import sample_forest_prior, sample_noise_prior
        Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
        self.train_data = (np.ascontiguousarray(X, dtype=np.float64), self.scaler(Y, train=True))
        bounds, feat_types = unpack_domain(self.domain)
        if self.prior_on_device:
            from .prior import sample_forest_prior_device
            self.forest = sample_forest_prior_device(self.num_trees, bounds, feat_types, self.alpha, self.beta,
                                                     self.num_samples, seed=int(self.sample_rng.integers(2**63)))
        else:
            self.forest = sample_forest_prior(self.num_trees, bounds, feat_types, self.alpha, self.beta, self.num_samples,
                                              self.sample_rng)
        self.noise = sample_noise_prior(self.gamma_prior_shape, self.gamma_prior_rate, self.num_samples, self.sample_rng)
        self.scale = np.ones((self.num_samples,))
        self._posterior = None
        return self
