"""Samples from the BARK prior (src/bark/fitting/bark_prior_sampler.py:15-93).

Two samplers with the same law: `sample_forest_prior` is host code driven by a numpy Generator exactly as the
reference's (so a reference script that passes its own `rng` keeps its stream semantics), and
`sample_forest_prior_device` grows all num_samples x m trees in one kernel launch (csrc/prior.cu, Philox streams
keyed by (seed, sample, tree)) and can leave the forest in HBM for the predictor.  The structure queries restate `get_node_subspace`
(src/bark/fitting/tree_traversal.py:49-86), `sample_splitting_rule` (src/bark/fitting/tree_proposals.py:78-97) and
`sample_binary_mask` (src/bark/utils/bit_operations.py:34-58) for a numpy Generator."""
from __future__ import annotations

import numpy as np

from .forest import FeatureTypeEnum, create_empty_forest

CAT, INT = FeatureTypeEnum.Cat.value, FeatureTypeEnum.Int.value


def _next_power_of_2(x: int) -> int:  # bit_operations.py:5-10
    return 1 if x <= 0 else 1 << int(x).bit_length()


def get_node_subspace(tree: np.ndarray, node_idx: int, bounds: np.ndarray, feat_types: np.ndarray) -> np.ndarray:
    sub = np.array(bounds, dtype=np.float64, copy=True)
    parent_idx = int(tree[node_idx]["parent"])
    while node_idx != 0:
        parent = tree[parent_idx]
        f = int(parent["feature_idx"])
        thr = float(parent["threshold"])
        if feat_types[f] == CAT:
            avail = int(sub[f, 1])
            if node_idx == int(parent["left"]):
                sub[f, 1] = int(thr) & avail
            else:
                sub[f, 1] = int(_next_power_of_2(avail) - 1 - thr) & avail
        elif node_idx == int(parent["left"]):
            sub[f, 1] = min(thr, sub[f, 1])
        else:
            sub[f, 0] = max(thr + (1 if feat_types[f] == INT else 0), sub[f, 0])
        node_idx, parent_idx = parent_idx, int(tree[parent_idx]["parent"])
    return sub


def sample_binary_mask(x: int, rng: np.random.Generator) -> int:
    n = bin(x).count("1")
    if n < 2:
        return 0
    sample = int(rng.integers(1, (1 << n) - 1))
    thr = 0
    for i in range(max(int(x).bit_length(), 0) + 1):
        if x & (1 << i):
            thr |= (sample & 1) << i
            sample >>= 1
    return thr


def sample_splitting_rule(subspace: np.ndarray, feat_types: np.ndarray, rng: np.random.Generator):
    f = int(rng.integers(0, subspace.shape[0]))
    if feat_types[f] == CAT:
        thr = sample_binary_mask(int(subspace[f, 1]), rng)
    elif feat_types[f] == INT:
        lo, hi = int(subspace[f, 0]), int(subspace[f, 1])
        thr = hi if lo == hi else int(rng.integers(lo, hi))
    else:
        thr = rng.uniform(subspace[f, 0], subspace[f, 1])
    return f, thr


def _sample_single_forest(m, bounds, feat_types, alpha, beta, rng):
    forest = create_empty_forest(m)
    for tree in forest:
        stack = [0]
        while stack:
            node = stack.pop()
            depth = int(tree[node]["depth"])
            if rng.uniform() > alpha * (1 + depth) ** (-beta):
                continue
            sub = get_node_subspace(tree, node, bounds, feat_types)
            f, thr = sample_splitting_rule(sub, feat_types, rng)
            if thr == 0 and feat_types[f] == CAT:
                continue
            if feat_types[f] == INT and thr == sub[f, 1]:
                continue
            free = np.flatnonzero(tree["active"] == 0)
            if free.size < 2:
                raise OverflowError("The tree container is not large enough")
            left, right = int(free[0]), int(free[1])
            parent = int(tree[node]["parent"])
            for child in (left, right):
                tree[child] = (1, 0, 0.0, 0, 0, node, depth + 1, 1)
            tree[node] = (0, f, np.float32(thr), left, right, parent, depth, 1)
            stack.extend([left, right])
    return forest


def sample_forest_prior(m, bounds, feat_types, alpha, beta, num_samples, rng: np.random.Generator | None = None):
    """(num_samples, m, L) forests drawn from the depth prior alpha (1+d)^-beta with uniform split rules."""
    rng = np.random.default_rng() if rng is None else rng
    bounds = np.asarray(bounds, dtype=np.float64)
    feat_types = np.asarray(feat_types)
    return np.array([_sample_single_forest(m, bounds, feat_types, alpha, beta, rng) for _ in range(num_samples)])


def sample_noise_prior(gamma_shape, gamma_rate, num_samples, rng: np.random.Generator | None = None):
    """bark_prior_sampler.py:87-93 (a Gamma(shape, rate) draw, as in the reference)."""
    rng = np.random.default_rng() if rng is None else rng
    return rng.gamma(shape=gamma_shape, scale=1 / gamma_rate, size=(num_samples,))


def sample_forest_prior_device(m, bounds, feat_types, alpha, beta, num_samples, seed=0, node_limit=100, device=None,
                               return_device=False):
    """(num_samples, m, L) forests from the same prior, grown on the GPU (csrc/prior.cu).  Returns the host structured
    array, or the `DeviceForest` itself with `return_device=True`."""
    import ctypes as C

    from . import _lib
    from .forest import DeviceForest, _as_device_f64, _feat_types_device, _ptr, _stream
    torch = _lib.require_cuda()
    dev = torch.device(device or "cuda")
    bounds = np.ascontiguousarray(bounds, dtype=np.float64)
    df = DeviceForest((num_samples, m, node_limit), dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    bd, fd = _as_device_f64(bounds, dev), _feat_types_device(feat_types, dev)
    _lib.check(_lib.load().bark_prior_sample(df.soa(), int(num_samples), int(m), int(node_limit), _ptr(bd), _ptr(fd),
                                             int(bounds.shape[0]), float(alpha), float(beta),
                                             C.c_uint64(int(seed) & (2**64 - 1)), _ptr(status), _stream()))
    if int(status.item()) & _lib.ST_TREE_OVERFLOW:
        raise OverflowError("The tree container is not large enough")  # tree_proposals.py:57-58
    return df if return_device else df.to_numpy()
