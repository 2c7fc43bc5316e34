"""CPU-side checks of the product package: C-ABI library loads and exports every symbol the header declares,
host logic, and 'no silent fallback' behaviour.  No GPU compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

import bark_b200
from bark_b200 import _build, _lib, domain, sampler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "bark_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bark_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bark_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.bark_abi_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.NodesSoA) == 8 * 8
    assert ctypes.sizeof(_lib.McmcDims) == 6 * 8
    assert ctypes.sizeof(_lib.Params) == 7 * 8 + 2 * 4
    assert bark_b200.NODE_RECORD_DTYPE.itemsize == 26


def test_workspace_sizes_and_argument_validation(lib):
    d = _lib.McmcDims(64, 2000, 10, 200, 100, 1600)
    nbytes = lib.bark_mcmc_workspace_bytes(ctypes.byref(d))
    assert 1e9 < nbytes < 8e9  # fits 180 GB HBM with room for 50x more chains
    bad = _lib.McmcDims(64, 2000, 10, 200, 100, 1601)  # p_cap not a multiple of 64
    assert lib.bark_mcmc_workspace_bytes(ctypes.byref(bad)) == 0
    assert lib.bark_mll_workspace_bytes(16, 250) > 0
    # invalid arguments are rejected before any CUDA call
    rc = lib.bark_gram_umma(None, None, 1, 4, 4, 3, 8, None, None, None, None, 1e-6, 0, None, None, None)
    assert rc == 1 and b"null" in lib.bark_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BarkError):
        bark_b200.pass_through_forest(bark_b200.create_empty_forest(2), np.zeros((3, 1)), [2])
    with pytest.raises(_lib.BarkError):
        p = bark_b200.BARKTrainParams(num_chains=1)
        bark_b200.run_bark_sampler((bark_b200.create_empty_forest(2)[None], [0.1], [1.0]),
                                   (np.zeros((3, 1)), np.zeros((3, 1))), (np.array([[0.0, 1.0]]), np.array([2])), p)


def test_product_package_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "bark_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_empty_forest_and_params():
    f = bark_b200.create_empty_forest(4)
    assert f.shape == (4, 100) and f[2, 0]["parent"] == 0xFFFFFFFF and f[2, 0]["active"] == 1 and f[2, 1]["active"] == 0
    p = bark_b200.BARKTrainParams()
    assert (p.warmup_steps, p.num_samples, p.steps_per_sample, p.num_chains) == (50, 5, 10, 1)
    c = p.to_c()
    assert list(c.proposal_weights) == [0.25, 0.25, 0.5] and c.use_softplus_transform == 1 and c.sample_scale == 0
    assert sampler.default_p_cap(200, 100) == 832 and sampler.default_p_cap(50, 100) == 256
    assert sampler.default_p_cap(2, 100) % 64 == 0


def test_status_to_exception_mapping():
    with pytest.raises(OverflowError, match="tree container"):
        sampler.raise_for_status(np.array([0, 1]))
    with pytest.raises(NotImplementedError):
        sampler.raise_for_status(np.array([2]))
    with pytest.raises(_lib.BarkError):
        sampler.raise_for_status(np.array([4]))
    sampler.raise_for_status(np.array([0, 0]))


def test_domain_adapters():
    dom = domain.Domain(inputs=domain.Inputs([
        domain.ContinuousInput("x0", (0.0, 1.0)), domain.DiscreteInput("i0", [1, 2, 3, 7]),
        domain.CategoricalInput("c0", ["a", "b", "c"])]))
    bounds, ft = domain.unpack_domain(dom)
    assert bounds.tolist() == [[0.0, 1.0], [1.0, 7.0], [0.0, 7.0]] and ft.tolist() == [2, 1, 0]
    b2, f2 = domain.unpack_domain((bounds, ft))
    assert np.array_equal(b2, bounds) and np.array_equal(f2, ft)
    assert domain.get_feature_bounds(dom.inputs.get()[2], "ordinal") == [0, 1, 2]


def test_mixture_moments_host():
    rng = np.random.default_rng(0)
    mu, var = rng.standard_normal((6, 9)), rng.random((6, 9))
    m, v = bark_b200.mixture_of_gaussians_as_normal(mu, var)
    assert np.allclose(m, mu.mean(0)) and np.allclose(v, (var + mu**2).mean(0) - mu.mean(0) ** 2)


# ------------------------------------------------------------ checkpoint format, prior sampler (host code)
def test_checkpoint_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    bounds = np.array([[0.0, 1.0], [0.0, 31.0], [2.0, 9.0]])
    ft = np.array([2, 0, 1])
    forest = bark_b200.sample_forest_prior(12, bounds, ft, 0.95, 2.0, 6, rng).reshape(2, 3, 12, 100)
    noise, scale = rng.random((2, 3)), rng.random((2, 3))
    X, y = rng.random((7, 3)), rng.random((7, 1))
    params = bark_b200.BARKTrainParams(warmup_steps=3, num_samples=3, num_chains=2, alpha=0.9)
    path = tmp_path / "samples.npz"
    bark_b200.save_samples(path, (forest, noise, scale), (X, y), params, extra={"note": "x"})
    ck = bark_b200.load_samples(path)
    assert ck["model"][0].dtype == bark_b200.NODE_RECORD_DTYPE and ck["model"][0].shape == forest.shape
    assert ck["model"][0].tobytes() == forest.tobytes()
    assert np.array_equal(ck["model"][1], noise) and np.array_equal(ck["model"][2], scale)
    assert np.array_equal(ck["data"][0], X) and np.array_equal(ck["data"][1], y)
    assert ck["params"].alpha == 0.9 and ck["params"].num_chains == 2 and ck["extra"] == {"note": "x"}
    assert np.allclose(ck["params"].proposal_weights, params.proposal_weights)
    with pytest.raises(TypeError):
        bark_b200.save_samples(path, (np.zeros((2, 3)), noise, scale))


def test_prior_sampler_structure_and_depth_law():
    """bark_prior_sampler.py:15-93: a node at depth d splits with probability alpha (1+d)^-beta; rules are drawn
    inside the node's subspace; the subspace restatement agrees with the oracle's."""
    from bark_b200 import prior
    from oracle import bark_oracle as O
    rng = np.random.default_rng(1)
    bounds = np.array([[0.0, 1.0], [-2.0, 3.0], [0.0, 31.0], [2.0, 9.0]])
    ft = np.array([2, 2, 0, 1])
    forests = bark_b200.sample_forest_prior(200, bounds, ft, 0.95, 2.0, 4, rng)
    assert forests.shape == (4, 200, 100) and forests.dtype == bark_b200.NODE_RECORD_DTYPE
    n_split_root = 0
    for tree in forests.reshape(-1, 100):
        act = np.flatnonzero(tree["active"])
        assert tree[0]["active"] and tree[0]["depth"] == 0
        n_split_root += int(tree[0]["is_leaf"] == 0)
        for i in act:
            nd = tree[i]
            sub = prior.get_node_subspace(tree, int(i), bounds, ft)
            want = O.get_node_subspace(tree, int(i), bounds, ft.astype(np.int64))
            assert np.array_equal(sub, want)
            if not nd["is_leaf"]:
                l, r, f = int(nd["left"]), int(nd["right"]), int(nd["feature_idx"])
                assert tree[l]["active"] and tree[r]["active"] and tree[l]["parent"] == i and tree[r]["parent"] == i
                assert tree[l]["depth"] == nd["depth"] + 1 == tree[r]["depth"]
                if ft[f] == 0:
                    assert 0 < int(nd["threshold"]) < int(sub[f, 1]) and int(nd["threshold"]) & ~int(sub[f, 1]) == 0
                elif ft[f] == 1:
                    assert sub[f, 0] <= nd["threshold"] < sub[f, 1]
                else:
                    assert np.float32(sub[f, 0]) <= nd["threshold"] <= np.float32(sub[f, 1])
    # root split probability alpha = 0.95 (every root rule is valid here): binomial(800, 0.95), 5 sigma
    assert abs(n_split_root / 800 - 0.95) < 5 * np.sqrt(0.95 * 0.05 / 800)
    assert all(prior._next_power_of_2(x) == O.next_power_of_2(x) for x in range(0, 70))
    noise = bark_b200.sample_noise_prior(1.5, 5.0, 20000, np.random.default_rng(2))
    assert abs(noise.mean() - 1.5 / 5.0) < 0.01


def test_new_entry_points_also_refuse_to_run_without_cuda():
    """The 8f rows obey the same rule as the hot path: no GPU, no result (never a host fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    f = bark_b200.create_empty_forest(3)[None]
    X, y = np.zeros((4, 2)), np.zeros((4, 1))
    dom = (np.array([[0.0, 1.0], [0.0, 1.0]]), np.array([2, 2]))
    with pytest.raises(_lib.BarkError):
        bark_b200.gp_sample_inverses((f, np.ones(1), np.ones(1)), (X, y + np.arange(4)[:, None]), dom)
    sur = bark_b200.BARKPriorSurrogate(dom, num_samples=2, num_trees=3, sample_seed=0).fit(X, y + np.arange(4)[:, None])
    assert sur.forest.shape == (2, 3, 100)  # drawing from the prior is host code, as in the reference
    with pytest.raises(_lib.BarkError):
        sur.predict(X)


def test_prior_surrogate_defaults_match_reference_data_model():
    """BARKPriorSurrogate keeps ITS OWN defaults (src/bofire_mixed/data_models/surrogates/bark.py:74-86):
    inverse-gamma(2.5, 9.0) noise prior, sample_seed 0 -> reproducible draws; the posterior surrogate keeps 1.5 / 5.0."""
    dom = (np.array([[0.0, 1.0], [0.0, 1.0]]), np.array([2, 2]))
    pr = bark_b200.BARKPriorSurrogate(dom)
    assert (pr.gamma_prior_shape, pr.gamma_prior_rate, pr.sample_seed, pr.num_samples) == (2.5, 9.0, 0, 5)
    assert (pr.alpha, pr.beta, pr.num_trees) == (0.95, 2.0, 50)
    po = bark_b200.BARKSurrogate(dom)
    assert (po.gamma_prior_shape, po.gamma_prior_rate) == (1.5, 5.0)
    X, y = np.random.default_rng(0).random((6, 2)), np.arange(6.0)[:, None]
    a = bark_b200.BARKPriorSurrogate(dom, num_trees=4).fit(X, y)
    b = bark_b200.BARKPriorSurrogate(dom, num_trees=4).fit(X, y)
    assert a.forest.tobytes() == b.forest.tobytes() and np.array_equal(a.noise, b.noise)


def test_distributed_sampler_rejects_more_ranks_than_chains_on_every_rank(monkeypatch):
    """`total < world` is checked before sharding, so no rank is left waiting in the all-gather."""
    from bark_b200 import distributed as D
    f = bark_b200.create_empty_forest(3)[None]
    for rank in (0, 1, 2):
        monkeypatch.setattr(D, "_world", lambda group=None, r=rank: (r, 3))
        with pytest.raises(ValueError, match="more ranks"):
            D.run_bark_sampler_distributed((np.tile(f, (2, 1, 1)), np.ones(2), np.ones(2)), (None, None), None,
                                           sampler.BARKTrainParams(num_chains=2), seed=1)


def test_bofire_compat_transform_and_validation():
    """The pandas / BoFire-shaped adapter: ORDINAL encoding of label-valued categoricals in feature order, the data
    model's validator (src/bofire_mixed/data_models/surrogates/bark.py:42-61), defaults of both data models."""
    import pandas as pd
    from bark_b200 import bofire_compat as BC
    from bark_b200.domain import CategoricalInput, ContinuousInput, DiscreteInput, Inputs
    inputs = Inputs([ContinuousInput("a", (0, 1)), CategoricalInput("c", ["u", "v", "w"]), DiscreteInput("k", [1, 2, 5])])
    df = pd.DataFrame({"k": [5, 1], "c": ["w", "u"], "a": [0.25, 0.5], "y": [1.0, 2.0]})
    X = BC.transform_inputs(inputs, df)
    assert X.dtype == np.float64 and np.array_equal(X, [[0.25, 2.0, 5.0], [0.5, 0.0, 1.0]])
    with pytest.raises(ValueError):
        BC.transform_inputs(inputs, pd.DataFrame({"a": [0.1], "c": ["zz"], "k": [1]}))
    dm = BC.BARKSurrogate(inputs=inputs, outputs=BC.Outputs())
    assert dm.input_preprocessing_specs == {"c": "ORDINAL"}
    assert (dm.warmup_steps, dm.num_samples, dm.steps_per_sample, dm.num_trees, dm.num_chains) == (50, 5, 10, 50, 1)
    assert (dm.gamma_prior_shape, dm.gamma_prior_rate, dm.grow_prune_weight, dm.change_weight) == (1.5, 5.0, 0.5, 1.0)
    with pytest.raises(ValueError, match="ordinal"):
        BC.BARKSurrogate(inputs=inputs, outputs=BC.Outputs(), input_preprocessing_specs={"c": "ONE_HOT"})
    pm = BC.BARKPriorSurrogate(inputs=inputs, outputs=BC.Outputs())
    assert (pm.gamma_prior_shape, pm.gamma_prior_rate, pm.sample_seed) == (2.5, 9.0, 0)
    sur = BC.surrogate_map(dm)
    assert sur.is_fitted is False and sur.bark_params.num_chains == 1
    w = sur.bark_params.proposal_weights
    assert np.allclose(w, [0.25, 0.25, 0.5])  # _bark_params_to_jitclass, surrogates/bark.py:24-36
    prior = BC.surrogate_map(pm).fit(df)  # drawing from the prior is host code; no GPU needed
    assert prior.forest.shape == (5, 50, 100) and prior.train_data[0].shape == (2, 3)


def test_predict_executed_op_count_follows_the_triangular_tiling():
    """`PosteriorState.umma_ops_per_candidate_sample` (what bench.py divides by the measured int8 peak): 7 digit planes x
    128-byte K tiles from the diagonal band down x column tiles of 192 (160 above 512 columns) -- csrc/predict_umma.cu
    pu_ntile / pu_kt_lo."""
    import types
    from bark_b200.predict import PosteriorState
    ops = PosteriorState.umma_ops_per_candidate_sample.fget

    def by_hand(k_pad, ntile):
        kt, total, nt = k_pad // 128, 0, 0
        while nt * ntile < k_pad:
            ncols = min(ntile, k_pad - nt * ntile)
            tiles = kt - (nt * ntile) // 128
            total += 2 * 7 * ncols * 128 * tiles
            nt += 1
        return total

    assert ops(types.SimpleNamespace(k_pad=None)) is None
    assert ops(types.SimpleNamespace(k_pad=128)) == 2 * 7 * 128 * 128
    assert ops(types.SimpleNamespace(k_pad=512)) == 2 * 7 * 128 * (192 * 4 + 192 * 3 + 128 * 1) == 2637824
    for k_pad in (256, 384, 512):
        assert ops(types.SimpleNamespace(k_pad=k_pad)) == by_hand(k_pad, 192)
    for k_pad in (640, 768):
        assert ops(types.SimpleNamespace(k_pad=k_pad)) == by_hand(k_pad, 160)
        assert ops(types.SimpleNamespace(k_pad=k_pad)) < 2 * 7 * k_pad * k_pad  # less than the full square
