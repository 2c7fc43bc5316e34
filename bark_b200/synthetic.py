"""Synthetic TreeFunction data (the generator BASELINE config 1 names): a forest sampled from the BARK depth
prior with N(0,1) leaf values, f(x) = sum_t leaf_value[t, leaf_t(x)]
(src/bofire_mixed/benchmarks/tree_function.py:19-88), bofire-free.  Tree evaluation runs on the GPU through
`pass_through_forest`."""
from __future__ import annotations

import numpy as np

from .forest import FeatureTypeEnum, create_empty_forest, pass_through_forest
from .surrogate import Standardize


def _grow_in_place(tree: np.ndarray, node: int, feature: int, threshold) -> None:
    """Host-side grow edit (src/bark/fitting/tree_proposals.py:146-165): first two inactive slots become leaves."""
    free = np.flatnonzero(tree["active"] == 0)
    if free.size < 2:
        raise OverflowError("The tree container is not large enough")
    left, right = int(free[0]), int(free[1])
    depth, parent = int(tree[node]["depth"]), int(tree[node]["parent"])
    for child in (left, right):
        tree[child] = (1, 0, 0.0, 0, 0, node, depth + 1, 1)
    tree[node] = (0, feature, np.float32(threshold), left, right, parent, depth, 1)


def sample_tree_structure_from_prior(m: int, n_features: int, rng: np.random.Generator, alpha=0.95, beta=2.0):
    """tree_function.py:36-57: split a node at depth d with probability alpha*(1+d)^-beta; feature uniform,
    threshold U(0,1) (also for categorical features, as in the reference)."""
    forest = create_empty_forest(m)
    for tree in forest:
        stack = [0]
        while stack:
            node = stack.pop()
            depth = int(tree[node]["depth"])
            if rng.uniform() > alpha * (1 + depth) ** (-beta):
                continue
            feature = int(rng.integers(n_features))
            threshold = rng.uniform(0, 1)
            _grow_in_place(tree, node, feature, threshold)
            stack.extend([int(tree[node]["left"]), int(tree[node]["right"])])
    return forest


class TreeFunction:
    def __init__(self, dim=5, cat_dim=0, num_cat=5, m=50, function_seed=1):
        self.dim, self.cat_dim, self.num_cat, self.m = dim, cat_dim, num_cat, m
        rng = np.random.default_rng(function_seed)
        self.forest = sample_tree_structure_from_prior(m, dim + cat_dim, rng)
        self.leaf_values = rng.standard_normal(self.forest.shape)
        self.feat_types = np.array([FeatureTypeEnum.Cont.value] * dim + [FeatureTypeEnum.Cat.value] * cat_dim, dtype=np.int64)
        self.bounds = np.array([[0.0, 1.0]] * dim + [[0.0, float((1 << num_cat) - 1)]] * cat_dim)

    def sample_inputs(self, n: int, rng: np.random.Generator) -> np.ndarray:
        cont = rng.uniform(size=(n, self.dim))
        cat = rng.integers(self.num_cat, size=(n, self.cat_dim)).astype(np.float64)
        return np.ascontiguousarray(np.hstack([cont, cat]))

    def __call__(self, X: np.ndarray) -> np.ndarray:
        leaves = pass_through_forest(self.forest, np.ascontiguousarray(X, dtype=np.float64), self.feat_types)
        return self.leaf_values[np.arange(self.m), leaves].sum(axis=1)


def synthetic_problem(n, dim=10, cat_dim=0, num_cat=5, m_true=50, seed=0, noise_std=0.1):
    """(X (n,D), y standardised (n,1), bounds (D,2), feat_types (D,), scaler): SURVEY section 8d workloads."""
    fn = TreeFunction(dim=dim, cat_dim=cat_dim, num_cat=num_cat, m=m_true, function_seed=1)
    rng = np.random.default_rng(seed)
    X = fn.sample_inputs(n, rng)
    y = fn(X) + noise_std * rng.standard_normal(n)
    scaler = Standardize()
    return X, scaler(y.reshape(-1, 1), train=True), fn.bounds.copy(), fn.feat_types.copy(), scaler
