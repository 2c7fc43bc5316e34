"""pandas / BoFire-shaped front end of the GPU surrogate, so that the reference's example scripts run on a swap of
imports.

What the reference's scripts touch (examples/regression/regression.py:47-119, examples/bayes_opt/*.py) is a small
slice of BoFire: a *data model* (`BARKSurrogate(inputs=..., outputs=..., **params)`,
src/bofire_mixed/data_models/surrogates/bark.py:15-86), `surrogate_map(data_model)` returning an object with
`fit(experiments: DataFrame)` / `predict(X: DataFrame) -> DataFrame[<out>_pred, <out>_sd]`, and
`inputs.transform(X, specs)` with ORDINAL encoding of categoricals (the validator of the data model insists on it,
:42-61).  This module restates that slice without bofire / pydantic and drives `bark_b200.BARKSurrogate`.  Real
bofire `Inputs` / `Outputs` objects are accepted too (duck typing: `.get()`, `.get_keys()`, feature class names).

    from bark_b200.bofire_compat import BARKSurrogate, surrogate_map      # instead of bofire_mixed.data_models...
    surrogate = surrogate_map(BARKSurrogate(inputs=domain.inputs, outputs=domain.outputs, num_chains=4))
    surrogate.fit(experiments)              # DataFrame with the input columns and the output column
    pred = surrogate.predict(test_x)        # DataFrame: y_pred, y_sd
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import surrogate as _sur
from .domain import Domain, Inputs, _kind


@dataclass
class ContinuousOutput:
    key: str = "y"


@dataclass
class Outputs:
    features: list = field(default_factory=lambda: [ContinuousOutput("y")])

    def get(self):
        return self.features

    def get_keys(self):
        return [f.key for f in self.features]


def _input_keys(inputs) -> list[str]:
    return [f.key for f in inputs.get()]


def transform_inputs(inputs, X):
    """`inputs.transform(X, specs)` with ORDINAL categoricals (the only encoding BARK accepts): a float matrix
    (N, D) in feature order; a category becomes its index in `feature.categories`; discrete / continuous columns
    pass through.  X: DataFrame (columns = feature keys) or an already-numeric array."""
    if not hasattr(X, "columns"):
        return np.ascontiguousarray(X, dtype=np.float64)
    cols = []
    for f in inputs.get():
        col = X[f.key]
        if _kind(f) == "CategoricalInput":
            lookup = {c: i for i, c in enumerate(f.categories)}
            vals = col.to_numpy()
            if vals.dtype.kind in "OUS":
                try:
                    vals = np.array([lookup[v] for v in vals], dtype=np.float64)
                except KeyError as exc:
                    raise ValueError(f"unknown category {exc} for feature {f.key}") from None
            cols.append(np.asarray(vals, dtype=np.float64))
        else:
            cols.append(col.to_numpy(dtype=np.float64))
    return np.ascontiguousarray(np.stack(cols, axis=1))


@dataclass
class BARKSurrogate:
    """Data model, field for field src/bofire_mixed/data_models/surrogates/bark.py:15-40."""
    inputs: Any = None
    outputs: Any = None
    warmup_steps: int = 50
    num_samples: int = 5
    steps_per_sample: int = 10
    alpha: float = 0.95
    beta: float = 2.0
    num_trees: int = 50
    use_softplus_transform: bool = True
    sample_scale: bool = False
    gamma_prior_shape: float = 1.5
    gamma_prior_rate: float = 5.0
    grow_prune_weight: float = 0.5
    change_weight: float = 1.0
    num_chains: int = 1
    verbose: bool = False
    input_preprocessing_specs: dict = field(default_factory=dict)
    type: str = "BARKSurrogate"

    def __post_init__(self):
        _validate_ordinal(self)


@dataclass
class BARKPriorSurrogate:
    """Data model of the prior-only surrogate, src/bofire_mixed/data_models/surrogates/bark.py:74-86."""
    inputs: Any = None
    outputs: Any = None
    alpha: float = 0.95
    beta: float = 2.0
    num_trees: int = 50
    gamma_prior_shape: float = 2.5
    gamma_prior_rate: float = 9.0
    sample_seed: int = 0
    num_samples: int = 5
    input_preprocessing_specs: dict = field(default_factory=dict)
    type: str = "BARKPriorSurrogate"

    def __post_init__(self):
        _validate_ordinal(self)


def _validate_ordinal(dm):
    """validate_input_preprocessing_specs (:42-61): categoricals must be ORDINAL-encoded; missing specs become ORDINAL."""
    if dm.inputs is None:
        return
    for f in dm.inputs.get():
        if _kind(f) == "CategoricalInput":
            enc = dm.input_preprocessing_specs.get(f.key, "ORDINAL")
            if str(getattr(enc, "name", enc)).upper() != "ORDINAL":
                raise ValueError("BARK based models have to use ordinal encoding for categoricals")
            dm.input_preprocessing_specs[f.key] = "ORDINAL"


class _FrameSurrogate:
    """`fit(experiments)` / `predict(X)` on DataFrames around a numpy-level surrogate (what bofire's Surrogate /
    TrainableSurrogate base classes do for src/bofire_mixed/surrogates/bark.py:71-94,123-149)."""

    def __init__(self, data_model, impl):
        self.data_model = data_model
        self.inputs = data_model.inputs
        self.outputs = data_model.outputs if data_model.outputs is not None else Outputs()
        self.input_preprocessing_specs = data_model.input_preprocessing_specs
        self._impl = impl

    # ---- attributes the strategies / optimiser read (surrogates/bark.py:54-69)
    def __getattr__(self, name):
        if name in ("forest", "noise", "scale", "train_data", "scaler", "bark_params", "is_fitted", "model_as_tuple",
                    "num_trees", "num_chains", "alpha", "beta", "save", "load"):
            return getattr(self._impl, name)
        raise AttributeError(name)

    def _fit(self, X, Y, **kwargs):
        Yn = Y.to_numpy(dtype=np.float64) if hasattr(Y, "to_numpy") else np.asarray(Y, dtype=np.float64)
        self._impl.fit(transform_inputs(self.inputs, X), Yn.reshape(-1, 1))

    def _predict(self, transformed_X, batched=False, predict_observed=True):
        return self._impl.predict(transform_inputs(self.inputs, transformed_X), batched=batched,
                                  predict_observed=predict_observed)

    def fit(self, experiments):
        """`experiments`: DataFrame with one column per input feature and the output column(s)."""
        out = self.outputs.get_keys()[0]
        valid = experiments[experiments[out].notna()] if hasattr(experiments[out], "notna") else experiments
        self._fit(valid[_input_keys(self.inputs)], valid[[out]])
        return self

    def predict(self, X):
        """DataFrame with `<out>_pred` and `<out>_sd` (bofire's Surrogate.predict contract)."""
        import pandas as pd
        mu, sd = self._predict(X[_input_keys(self.inputs)])
        out = self.outputs.get_keys()[0]
        return pd.DataFrame({f"{out}_pred": mu[:, 0], f"{out}_sd": sd[:, 0]}, index=getattr(X, "index", None))


def surrogate_map(data_model, seed=None):
    """src/bofire_mixed/data_models/surrogates/mapper.py: data model -> surrogate object (GPU-backed)."""
    domain = Domain(inputs=data_model.inputs, outputs=data_model.outputs)
    if data_model.type == "BARKPriorSurrogate":
        impl = _sur.BARKPriorSurrogate(domain, num_samples=data_model.num_samples, sample_seed=data_model.sample_seed,
                                       alpha=data_model.alpha, beta=data_model.beta, num_trees=data_model.num_trees,
                                       gamma_prior_shape=data_model.gamma_prior_shape,
                                       gamma_prior_rate=data_model.gamma_prior_rate)
    elif data_model.type == "BARKSurrogate":
        impl = _sur.BARKSurrogate(
            domain, warmup_steps=data_model.warmup_steps, num_samples=data_model.num_samples,
            steps_per_sample=data_model.steps_per_sample, alpha=data_model.alpha, beta=data_model.beta,
            num_trees=data_model.num_trees, use_softplus_transform=data_model.use_softplus_transform,
            sample_scale=data_model.sample_scale, gamma_prior_shape=data_model.gamma_prior_shape,
            gamma_prior_rate=data_model.gamma_prior_rate, grow_prune_weight=data_model.grow_prune_weight,
            change_weight=data_model.change_weight, num_chains=data_model.num_chains, verbose=data_model.verbose, seed=seed)
    else:
        raise KeyError(f"no GPU surrogate for data model type {data_model.type!r}")
    return _FrameSurrogate(data_model, impl)


__all__ = ["BARKSurrogate", "BARKPriorSurrogate", "surrogate_map", "transform_inputs", "Outputs", "ContinuousOutput",
           "Inputs", "Domain"]
