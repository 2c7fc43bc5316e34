"""bofire-free stand-ins for the tiny part of the BoFire domain API the hot path touches.

Mirrors `src/bofire_mixed/domain.py:27-65`: `get_feature_bounds(feature, "bitmask")` and
`get_feature_types_array(domain)`.  Real bofire feature objects are accepted too (matched by class name), so
the reference's `BARKSurrogate` can hand its `Domain` straight to `run_bark_sampler`."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from .forest import FeatureTypeEnum


@dataclass
class ContinuousInput:
    key: str
    bounds: tuple[float, float]


@dataclass
class DiscreteInput:
    key: str
    values: Sequence[float]

    @property
    def lower_bound(self):
        return min(self.values)

    @property
    def upper_bound(self):
        return max(self.values)


@dataclass
class CategoricalInput:
    key: str
    categories: Sequence[str]


@dataclass
class Inputs:
    features: list = field(default_factory=list)

    def get(self):
        return self.features

    def __len__(self):
        return len(self.features)


@dataclass
class Domain:
    inputs: Inputs
    outputs: object = None


def _kind(feature) -> str:
    return type(feature).__name__


def get_feature_bounds(feature, encoding=None):
    """src/bofire_mixed/domain.py:27-43."""
    k = _kind(feature)
    if k == "CategoricalInput":
        cats = feature.categories
        if encoding == "bitmask":
            return (0, (1 << len(cats)) - 1)
        if encoding == "ordinal":
            return list(range(len(cats)))
        return cats
    if k == "DiscreteInput":
        return (feature.lower_bound, feature.upper_bound)
    if k == "ContinuousInput":
        return tuple(feature.bounds)
    raise TypeError(f"Cannot get bounds for feature of type {k}")


def get_feature_types_array(domain) -> np.ndarray:
    """src/bofire_mixed/domain.py:55-65."""
    out = []
    for feat in domain.inputs.get():
        k = _kind(feat)
        out.append(FeatureTypeEnum.Cat.value if k == "CategoricalInput"
                   else FeatureTypeEnum.Int.value if k == "DiscreteInput" else FeatureTypeEnum.Cont.value)
    return np.array(out)


def unpack_domain(domain):
    """(bounds (D,2) f64, feat_types (D,) int64) from a Domain-like object or a `(bounds, feat_types)` pair."""
    if isinstance(domain, (tuple, list)) and len(domain) == 2 and not hasattr(domain, "inputs"):
        bounds, ft = domain
        return np.ascontiguousarray(bounds, dtype=np.float64), np.ascontiguousarray(ft, dtype=np.int64)
    bounds = np.array([get_feature_bounds(f, encoding="bitmask") for f in domain.inputs.get()], dtype=np.float64)
    return bounds, get_feature_types_array(domain).astype(np.int64)
