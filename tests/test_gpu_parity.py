"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and the committed golden
fixtures.  Run with `-m gpu` on a B200."""
import os

import numpy as np
import pytest

import bark_b200 as B
from bark_b200 import sampler as S
from oracle import bark_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def rec(a):
    return np.ascontiguousarray(a).view(O.NODE_RECORD_DTYPE).reshape(a.shape[:-1] + (a.shape[-1] // 26,))


def load(name):
    return np.load(os.path.join(G, name))


def random_forests(n_forests, m, bounds, ft, sweeps=30, seed=0):
    """Forests with realistic structure: run the oracle's prior-like proposal chain (accept everything valid)."""
    rng = np.random.default_rng(seed)
    cdf = np.cumsum([0.5, 0.1, 0.4])
    out = np.tile(O.create_empty_forest(m), (n_forests, 1, 1))
    for f in out:
        for _ in range(sweeps):
            for t in range(m):
                tape = rng.random(5)
                new, lqp, st = O.get_tree_proposal(f[t], bounds, ft, 0.95, 2.0, cdf, tape, 0)
                if np.isfinite(lqp):
                    f[t] = new
    return out


# ----------------------------------------------------------------------------------------- a1 codec
def test_codec_roundtrip_lossless():
    import torch
    from bark_b200.forest import DeviceForest
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, size=(3, 7, 100, 26), dtype=np.uint8)  # arbitrary bytes incl. NaN thresholds
    nodes = raw.view(O.NODE_RECORD_DTYPE).reshape(3, 7, 100)
    back = DeviceForest.from_numpy(nodes).to_numpy()
    assert back.tobytes() == nodes.tobytes()
    df = DeviceForest.from_numpy(nodes)
    assert np.array_equal(df.parent.cpu().numpy().view(np.uint32).reshape(3, 7, 100), nodes["parent"])
    assert np.array_equal(df.threshold.cpu().numpy().view(np.uint32).reshape(3, 7, 100), nodes["threshold"].view(np.uint32))
    empty = B.create_empty_forest(5)
    assert DeviceForest.from_numpy(empty).to_numpy().tobytes() == empty.tobytes()
    torch.cuda.synchronize()


# ----------------------------------------------------------------------------------------- a2 traversal
def test_traversal_kat():
    k = load("kat.npz")
    ft = np.array([2])
    assert B.pass_through_forest(rec(k["kat_tree"]).reshape(1, -1), k["x20"], ft)[:, 0].tolist() == [3] * 5 + [4] * 5 + [2] * 10
    assert B.pass_through_forest(rec(k["t_thr"]).reshape(1, -1), k["xb"], ft)[:, 0].tolist() == [1, 1, 2, 1]
    assert B.pass_through_forest(rec(k["t_cat"]).reshape(1, -1), k["xc"], np.array([0]))[:, 0].tolist() == [2, 1, 1, 2, 1]


@pytest.mark.parametrize("tag", ["cont", "mixed"])
def test_traversal_golden(tag):
    s, g = load(f"sampler_{tag}.npz"), load(f"functions_{tag}.npz")
    ns = rec(s["node_samples"])
    forests = ns.reshape(-1, *ns.shape[-2:])
    got = B.pass_through_forest(forests, g["Xp"], s["feat_types"])
    assert got.dtype == np.uint32 and np.array_equal(got, g["leaves"])  # bit-exact vs the reference's own output


@pytest.mark.parametrize("n,m,dims", [(1, 1, (1, 0)), (257, 33, (3, 2)), (2000, 200, (10, 0)), (500, 100, (6, 4))])
def test_traversal_random_vs_oracle(n, m, dims):
    dim, cat = dims
    fn = O.TreeFunction(dim=dim, cat_dim=cat, num_cat=5, m=2, function_seed=3)
    forests = random_forests(2, m, fn.bounds, fn.feat_types, sweeps=12, seed=n)
    X = fn.sample_inputs(n, np.random.default_rng(n))
    # put points exactly on, just above and just below f32 thresholds; signed zeros
    f0 = forests[0]
    k = 0
    for t in range(m):
        for nd in f0[t]:
            if nd["active"] and not nd["is_leaf"] and fn.feat_types[nd["feature_idx"]] == 2 and k + 3 <= n:
                thr = float(nd["threshold"])
                X[k, nd["feature_idx"]], X[k + 1, nd["feature_idx"]], X[k + 2, nd["feature_idx"]] = \
                    thr, np.nextafter(thr, 2.0), np.nextafter(thr, -2.0)
                k += 3
    if n > 3:
        X[-1, 0] = -0.0
    want = np.stack([O.pass_through_forest(f, X, fn.feat_types) for f in forests])
    got = B.pass_through_forest(forests, X, fn.feat_types)
    assert np.array_equal(got, want)


def test_traversal_empty_inputs():
    f = B.create_empty_forest(3)
    assert B.pass_through_forest(f, np.zeros((0, 2)), [2, 2]).shape == (0, 3)
    out = B.pass_through_forest(f, np.zeros((5, 2)), [2, 2])
    assert out.shape == (5, 3) and not out.any()  # root-only trees: every point in slot 0


# ----------------------------------------------------------------------------------------- a3/a4 Gram
@pytest.mark.parametrize("tag", ["cont", "mixed"])
def test_gram_golden_bit_exact(tag):
    s, g = load(f"sampler_{tag}.npz"), load(f"functions_{tag}.npz")
    ns = rec(s["node_samples"])
    forests = ns.reshape(-1, *ns.shape[-2:])
    X, ft = s["X"], s["feat_types"]
    assert np.array_equal(B.batched_forest_gram_matrix(forests, X, X, ft), g["gram"])
    assert np.array_equal(B.batched_forest_gram_matrix(forests, g["Xp"], X, ft), g["gram_cross"])
    assert np.allclose(B.batched_forest_gram_matrix_no_null(forests, X, X, ft), g["gram_nonull"], rtol=1e-15, atol=0)
    k = load("kat.npz")
    K = B.forest_gram_matrix(rec(k["kat_tree"]).reshape(1, -1), k["x20"], k["x20"], np.array([2]))
    assert np.array_equal(K, k["kat_K"]) and K.sum() == 150
    for i in range(6):
        assert np.array_equal(B.get_leaf_vectors(forests[0][i], X, ft), g[f"leafvec{i}"])


@pytest.mark.parametrize("n,n2,m", [(65, 1, 3), (300, 129, 51), (700, 700, 201)])
def test_gram_counts_random(n, n2, m):
    fn = O.TreeFunction(dim=4, cat_dim=2, num_cat=5, m=2, function_seed=3)
    forests = random_forests(2, m, fn.bounds, fn.feat_types, sweeps=10, seed=m)
    rng = np.random.default_rng(m)
    X1, X2 = fn.sample_inputs(n, rng), fn.sample_inputs(n2, rng)
    got = B.forest_gram_counts(forests, X1, X2, fn.feat_types)
    want = np.stack([O.forest_gram_counts(f, X1, X2, fn.feat_types) for f in forests])
    assert got.dtype == np.int32 and np.array_equal(got, want)
    # K0 = (1/m)*count: multiply by the rounded reciprocal (differs from count/m in the last ulp)
    K0 = B.batched_forest_gram_matrix(forests, X1, X2, fn.feat_types)
    assert np.array_equal(K0, (1 / m) * want.astype(np.float64))


def test_gram_kernel_matrix_bit_exact():
    import torch
    from bark_b200.forest import gram_to_kernel_device
    rng = np.random.default_rng(0)
    m = 51
    cnt = rng.integers(0, m + 1, size=(3, 40, 40)).astype(np.int32)
    scale, noise = rng.random(3) + 0.5, rng.random(3) * 0.3
    K = gram_to_kernel_device(torch.from_numpy(cnt).cuda(), m, torch.from_numpy(scale).cuda(), torch.from_numpy(noise).cuda())
    want = np.stack([scale[i] * ((1 / m) * cnt[i].astype(np.float64)) + (1e-6 + noise[i]) * np.eye(40) for i in range(3)])
    assert np.array_equal(K.cpu().numpy(), want)


# ----------------------------------------------------------------------------------------- a5 MLL
@pytest.mark.parametrize("tag", ["cont", "mixed"])
def test_mll_golden(tag):
    s, g = load(f"sampler_{tag}.npz"), load(f"functions_{tag}.npz")
    ns = rec(s["node_samples"])
    forests = ns.reshape(-1, *ns.shape[-2:])
    got = B.forest_mll(forests, s["noise_samples"].reshape(-1), s["scale_samples"].reshape(-1), s["X"], s["y"], s["feat_types"])
    assert np.allclose(got, g["mll"], rtol=1e-9, atol=0)  # north_star tolerance: 1e-9 relative
    assert np.abs(got / g["mll"] - 1).max() < 1e-11


@pytest.mark.parametrize("n", [1, 63, 64, 65, 250, 517])
def test_mll_batched_random_spd(n):
    import torch
    from bark_b200.mll import mll_batched_device
    rng = np.random.default_rng(n)
    b = 5
    Z = (rng.random((b, n, max(n // 3, 2))) < 0.3).astype(np.float64)
    K = np.einsum("bik,bjk->bij", Z, Z) / 7 + (0.05 + rng.random((b, 1, 1))) * np.eye(n)
    y = rng.standard_normal(n)
    val, ld, quad, st = mll_batched_device(torch.from_numpy(K.copy()).cuda(), torch.from_numpy(y).cuda())
    assert int(st.max()) == 0
    for i in range(b):
        want_ld = np.linalg.slogdet(K[i])[1]
        want_q = y @ np.linalg.solve(K[i], y)
        assert ld[i].item() == pytest.approx(want_ld, rel=1e-11, abs=1e-11)
        assert quad[i].item() == pytest.approx(want_q, rel=1e-11)
        assert val[i].item() == pytest.approx(O.mll_cholesky(K[i], y), rel=1e-10, abs=1e-11)


def test_mll_flags_non_spd():
    import torch
    from bark_b200.mll import mll_batched_device
    K = -np.eye(8)[None]
    _, _, _, st = mll_batched_device(torch.from_numpy(K.copy()).cuda(), torch.zeros(8, dtype=torch.float64).cuda())
    assert int(st[0]) & 8


# ----------------------------------------------------------------------------------------- a8-a13 MCMC replay
def replay_case(n, dim, cat, m, chains, warm, ns, sps, seed, **pkw):
    X, y, bounds, ft, _ = O.synthetic_problem(n, dim=dim, cat_dim=cat, num_cat=4, m_true=10, seed=seed)
    p = O.BARKTrainParams(warmup_steps=warm, num_samples=ns, steps_per_sample=sps, num_chains=chains, **pkw)
    sweeps = warm + ns * sps
    tape = O.make_tape(np.random.default_rng(seed + 100), chains, sweeps, m)
    trace_o = np.zeros((chains, sweeps, m + 1, 3))
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    noise0, scale0 = np.full(chains, 0.1), np.full(chains, 1.0)
    want = O.run_bark_sampler((f0.copy(), noise0, scale0), (X, y), bounds, ft, p, tape=tape, trace=trace_o)
    pg = B.BARKTrainParams(warmup_steps=warm, num_samples=ns, steps_per_sample=sps, num_chains=chains, **pkw)
    got = B.run_bark_sampler((f0.copy(), noise0, scale0), (X, y), (bounds, ft), pg, tape=tape, return_trace=True,
                             return_info=True)
    return want, trace_o, got, (X, y, bounds, ft)


@pytest.mark.parametrize("case", [
    dict(n=50, dim=5, cat=0, m=50, chains=1, warm=10, ns=2, sps=5, seed=0),      # BASELINE config 1 shape
    dict(n=60, dim=3, cat=2, m=12, chains=3, warm=8, ns=2, sps=4, seed=1),       # mixed categorical
    dict(n=130, dim=4, cat=0, m=20, chains=2, warm=6, ns=1, sps=4, seed=2),
    dict(n=40, dim=2, cat=1, m=8, chains=2, warm=5, ns=2, sps=3, seed=3, sample_scale=True),
    dict(n=40, dim=2, cat=1, m=8, chains=2, warm=5, ns=2, sps=3, seed=4, sample_scale=True, use_softplus_transform=False),
])
def test_mcmc_replay_matches_oracle(case):
    """Same pre-drawn random numbers -> same trajectory: every proposal's log q/prior ratio, proposed log-MLL and
    accept bit, and the sampled forests byte for byte."""
    want, trace_o, got, _ = replay_case(**case)
    ns_g, noise_g, scale_g, trace_g, info = got
    lq_o, lq_g = trace_o[..., 0], trace_g[..., 0]
    assert np.array_equal(np.isfinite(lq_o), np.isfinite(lq_g))
    fin = np.isfinite(lq_o)
    tree = fin.copy(); tree[..., -1] = False
    assert np.allclose(lq_g[tree], lq_o[tree], rtol=0, atol=1e-12)
    # hyper step: with sample_scale the reference's log_q divides a squared 1e-8-sized difference by 1e-16
    # (noise_scale_proposals.py:113-121), which amplifies last-ulp libm differences to ~1e-8
    hyp_tol = 1e-6 if case.get("sample_scale") else 1e-12
    assert np.allclose(lq_g[..., -1], lq_o[..., -1], rtol=hyp_tol, atol=hyp_tol)
    flips = trace_o[..., 2] != trace_g[..., 2]
    assert not flips.any(), f"{flips.sum()} accept decisions differ"
    # proposed log-MLL of every valid proposal: 1e-9 relative (north_star)
    rel = np.abs(trace_g[..., 1][fin] - trace_o[..., 1][fin]) / np.maximum(np.abs(trace_o[..., 1][fin]), 1e-300)
    assert rel.max() < 1e-9, rel.max()
    assert ns_g.tobytes() == want[0].tobytes()
    assert np.allclose(noise_g, want[1], rtol=1e-12, atol=0) and np.allclose(scale_g, want[2], rtol=1e-12, atol=0)
    assert int(info["tree_proposals"].sum()) == trace_o.shape[0] * trace_o.shape[1] * (trace_o.shape[2] - 1)
    assert int(info["accepted"].sum()) == int(trace_o[:, :, :-1, 2].sum())
    assert int(info["hyper_accepted"].sum()) == int(trace_o[:, :, -1, 2].sum())


def test_mcmc_state_consistency_after_sweeps():
    """Leaf-space state kept by the kernels == state recomputed from the final forest by the oracle."""
    X, y, bounds, ft, _ = O.synthetic_problem(90, dim=3, cat_dim=1, num_cat=4, m_true=10, seed=5)
    chains, m = 2, 16
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    st = S.ChainState(f0, np.full(chains, 0.1), np.full(chains, 1.0), X, y, bounds, ft)
    p = B.BARKTrainParams(num_chains=chains)
    st.sweeps(p, 25, seed=123)
    r = st.read()
    assert int(r["status"].max()) == 0
    forest = st.dforest.to_numpy()
    for c in range(chains):
        ex = st.export(c)
        leaves = O.pass_through_forest(forest[c], X, ft)
        cm = ex["colmap"]
        P = st.p_cap
        Z = np.zeros((X.shape[0], P))
        for t in range(m):
            for sl in range(100):
                is_leaf = forest[c, t, sl]["active"] and forest[c, t, sl]["is_leaf"]
                assert (cm[t, sl] >= 0) == bool(is_leaf)
            Z[np.arange(X.shape[0]), cm[t, leaves[:, t]]] = 1
        assert np.array_equal(ex["A"], (Z.T @ Z).astype(np.int32))
        bits = np.unpackbits(ex["bits"].view(np.uint8), axis=1, bitorder="little")[:, :X.shape[0]]
        assert np.array_equal(bits, Z.T.astype(np.uint8))
        noise, scale = r["noise"][c].item(), r["scale"][c].item()
        cc = (noise + 1e-6) * m / scale
        Bm = cc * np.eye(P) + Z.T @ Z
        assert np.abs(ex["Binv"] @ Bm - np.eye(P)).max() < 1e-9
        K = O.kernel_matrix(forest[c], X, ft, noise, scale)
        assert r["mll"][c].item() == pytest.approx(O.mll_from_kernel(K, y), rel=1e-9)
        assert int(r["p_used"][c]) == int((forest[c]["active"] & forest[c]["is_leaf"]).sum())


def test_mcmc_philox_run_statistics():
    """Free-running (Philox) chains: valid structure, finite MLL equal to the refactorised one, sane acceptance."""
    X, y, bounds, ft, _ = O.synthetic_problem(120, dim=4, cat_dim=0, m_true=10, seed=7)
    chains, m = 8, 20
    p = B.BARKTrainParams(warmup_steps=30, num_samples=3, steps_per_sample=5, num_chains=chains)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    ns, noise, scale, info = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p,
                                                seed=42, return_info=True)
    assert ns.shape == (chains, 3, m, 100) and noise.shape == (chains, 3)
    ns2, noise2, _, _ = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p,
                                           seed=42, return_info=True)
    assert ns.tobytes() == ns2.tobytes() and np.array_equal(noise, noise2)  # deterministic given the seed
    acc = info["accepted"].sum() / info["tree_proposals"].sum()
    assert 0.05 < acc < 0.9
    assert len({ns[c].tobytes() for c in range(chains)}) == chains  # chains use distinct streams
    for c in range(chains):
        K = O.kernel_matrix(ns[c, -1], X, ft, noise[c, -1], scale[c, -1])
        assert info["mll"][c] == pytest.approx(O.mll_from_kernel(K, y), rel=1e-9)
        for t in range(m):  # structural invariants of every sampled tree
            tr = ns[c, -1, t]
            for i in np.flatnonzero(tr["active"] & (1 - tr["is_leaf"])):
                assert tr[tr[i]["left"]]["active"] and tr[tr[i]["right"]]["active"]
                assert tr[tr[i]["left"]]["parent"] == i and tr[tr[i]["left"]]["depth"] == tr[i]["depth"] + 1


def test_mcmc_error_conventions():
    X, y, bounds, ft, _ = O.synthetic_problem(30, dim=2, m_true=5, seed=1)
    f0 = np.tile(O.create_empty_forest(4), (1, 1, 1))
    p = B.BARKTrainParams(warmup_steps=1, num_samples=1, steps_per_sample=1, num_chains=1,
                          use_softplus_transform=False, sample_scale=False)
    with pytest.raises(NotImplementedError):
        B.run_bark_sampler((f0, [0.1], [1.0]), (X, y), (bounds, ft), p, seed=1)
    # tree container overflow: node_limit 3 cannot hold a second grow
    small = np.zeros((1, 4, 3), dtype=O.NODE_RECORD_DTYPE)
    small[:, :, 0] = (1, 0, 0, 0, 0, 0xFFFFFFFF, 0, 1)
    p2 = B.BARKTrainParams(warmup_steps=40, num_samples=1, steps_per_sample=1, num_chains=1, proposal_weights=(1.0, 0.0, 0.0))
    with pytest.raises(OverflowError):
        B.run_bark_sampler((small, [0.1], [1.0]), (X, y), (bounds, ft), p2, seed=1)


# ----------------------------------------------------------------------------------------- a14-a15 predict
@pytest.mark.parametrize("tag", ["cont", "mixed"])
def test_predict_golden(tag):
    s, g = load(f"sampler_{tag}.npz"), load(f"functions_{tag}.npz")
    model = (rec(s["node_samples"]), s["noise_samples"], s["scale_samples"])
    mu, var = B.forest_predict(model, (s["X"], s["y"]), g["Xp"], (s["bounds"], s["feat_types"]))
    assert mu.shape == g["pred_mu"].shape
    assert np.allclose(mu, g["pred_mu"], rtol=1e-9, atol=1e-11)
    assert np.allclose(var, g["pred_var"], rtol=1e-8, atol=1e-11)
    mm, mv = B.mixture_of_gaussians_as_normal(mu, var)
    assert np.allclose(mm, g["mix_mu"], rtol=1e-9, atol=1e-11) and np.allclose(mv, g["mix_var"], rtol=1e-8, atol=1e-11)


def test_surrogate_fit_predict_end_to_end():
    X, y, bounds, ft, scaler = O.synthetic_problem(80, dim=3, cat_dim=1, num_cat=4, m_true=10, seed=9)
    y_raw = y * 2.5 + 1.0
    sur = B.BARKSurrogate((bounds, ft), warmup_steps=20, num_samples=3, steps_per_sample=4, num_trees=10, num_chains=2, seed=5)
    sur.fit(X, y_raw)
    assert sur.forest.shape == (2, 3, 10, 100) and sur.is_fitted
    Xc = X[:17]
    mu, sd = sur.predict(Xc)
    mu_b, sd_b = sur.predict(Xc, batched=True)
    assert mu.shape == (17, 1) and mu_b.shape == (6, 17, 1)
    sc = O.Standardize(); sc.mean, sc.std = sur.scaler.mean, sur.scaler.std
    want_mu, want_sd = O.surrogate_predict(sur.model_as_tuple(), sur.train_data, Xc, ft, sc)
    assert np.allclose(mu, want_mu, rtol=1e-9, atol=1e-10) and np.allclose(sd, want_sd, rtol=1e-8, atol=1e-10)
    want_mu_b, want_sd_b = O.surrogate_predict(sur.model_as_tuple(), sur.train_data, Xc, ft, sc, batched=True)
    assert np.allclose(mu_b, want_mu_b, rtol=1e-9, atol=1e-10) and np.allclose(sd_b, want_sd_b, rtol=1e-8, atol=1e-10)
    sur.fit(X, y_raw)  # warm start: continues from the last sample with warmup 0 (surrogates/bark.py:131-141)
    assert sur.bark_params.warmup_steps == 0 and sur.forest.shape == (2, 3, 10, 100)


def test_synthetic_generator_matches_oracle():
    from bark_b200 import synthetic
    for kw in (dict(n=70, dim=5), dict(n=90, dim=6, cat_dim=4, num_cat=5, m_true=20, seed=3)):
        a, b = synthetic.synthetic_problem(**kw), O.synthetic_problem(**kw)
        for x, y in zip(a[:4], b[:4]):
            assert np.array_equal(x, y)
    assert synthetic.TreeFunction(m=7).forest.tobytes() == O.TreeFunction(m=7).forest.tobytes()


@pytest.mark.parametrize("n,n2,m", [(65, 1, 3), (300, 129, 51), (700, 700, 201), (128, 256, 64)])
def test_gram_umma_tensor_core_counts(n, n2, m):
    """int8 one-hot tcgen05 GEMM: exact counts and bit-exact FP64 kernel matrix vs the oracle."""
    import torch
    from bark_b200.forest import DeviceForest, _as_device_f64, _feat_types_device, gram_umma_device, traverse_device
    fn = O.TreeFunction(dim=4, cat_dim=2, num_cat=5, m=2, function_seed=3)
    forests = random_forests(2, m, fn.bounds, fn.feat_types, sweeps=10, seed=m)
    rng = np.random.default_rng(m)
    X1, X2 = fn.sample_inputs(n, rng), fn.sample_inputs(n2, rng)
    dev = torch.device("cuda")
    df = DeviceForest.from_numpy(forests, dev)
    ft = _feat_types_device(fn.feat_types, dev)
    la = traverse_device(df, _as_device_f64(X1, dev), ft)
    lb = traverse_device(df, _as_device_f64(X2, dev), ft)
    want = np.stack([O.forest_gram_counts(f, X1, X2, fn.feat_types) for f in forests])
    scale = torch.tensor([0.7, 1.3], dtype=torch.float64, device=dev)
    cnt, K = gram_umma_device(la, lb, scale=scale)
    assert np.array_equal(cnt.cpu().numpy(), want)
    assert np.array_equal(K.cpu().numpy(), scale.cpu().numpy()[:, None, None] * ((1 / m) * want.astype(np.float64)))
    # square case with the noise diagonal, operands shared (Z built once)
    noise = torch.tensor([0.1, 0.02], dtype=torch.float64, device=dev)
    cnt2, K2 = gram_umma_device(la, la, scale=scale, noise=noise)
    want2 = np.stack([O.forest_gram_counts(f, X1, X1, fn.feat_types) for f in forests])
    assert np.array_equal(cnt2.cpu().numpy(), want2)
    wantK = np.stack([scale[i].item() * ((1 / m) * want2[i].astype(np.float64)) + (1e-6 + noise[i].item()) * np.eye(n)
                      for i in range(2)])
    assert np.array_equal(K2.cpu().numpy(), wantK)


@pytest.mark.parametrize("n,n2,m,slots", [(200, 130, 64, 40), (129, 129, 7, 256), (33, 500, 300, 3)])
def test_gram_umma_wide_leaf_alphabets(n, n2, m, slots):
    """Leaf ids drawn directly from [0, slots): more occupied (tree, slot) columns than one pass of the operand build
    holds (20 K tiles at m = 64, slots = 40), presence masks of 1..8 words per tree, ragged tiles; counts stay exact, and an
    id >= slots is refused."""
    import torch
    from bark_b200.forest import gram_umma_device
    rng = np.random.default_rng(slots)
    la = rng.integers(0, slots, size=(2, n, m), dtype=np.int64)
    lb = rng.integers(0, slots, size=(2, n2, m), dtype=np.int64)
    want = (la[:, :, None, :] == lb[:, None, :, :]).sum(-1).astype(np.int32)
    dev = torch.device("cuda")
    ta = torch.tensor(la.astype(np.int32), device=dev)
    tb = torch.tensor(lb.astype(np.int32), device=dev)
    cnt, _ = gram_umma_device(ta, tb, slots=slots)
    assert np.array_equal(cnt.cpu().numpy(), want)
    cnt2, _ = gram_umma_device(ta, ta, slots=slots)
    assert np.array_equal(cnt2.cpu().numpy(), (la[:, :, None, :] == la[:, None, :, :]).sum(-1).astype(np.int32))
    bad = ta.clone()
    bad[1, n // 2, m // 2] = slots
    with pytest.raises(B.BarkError):
        gram_umma_device(bad, tb, slots=slots)


def test_predict_both_paths_and_edges():
    """tcgen05 int8-sliced variance and the FP64 gather kernel agree with the oracle; ragged / empty candidate sets."""
    import torch
    s, g = load("sampler_mixed.npz"), load("functions_mixed.npz")
    model = (rec(s["node_samples"]), s["noise_samples"], s["scale_samples"])
    X, y, ft = s["X"], s["y"], s["feat_types"]
    rng = np.random.default_rng(3)
    cand = np.vstack([g["Xp"], g["Xp"][rng.integers(len(g["Xp"]), size=200)]])  # 200+ candidates: > one 128-row tile
    want_mu, want_var = O.forest_predict(model, (X, y), cand, ft)
    for tc in (True, False):
        ps = B.PosteriorState(model, (X, y), ft, cand.shape[1], tensor_cores=tc)
        assert (ps.prep is not None) == tc
        mu, var = ps.predict_device(torch.from_numpy(cand).cuda(), mode=0)
        assert np.allclose(mu.cpu().numpy(), want_mu, rtol=1e-9, atol=1e-11)
        assert np.allclose(var.cpu().numpy(), want_var, rtol=1e-8, atol=1e-11)
        m1, v1 = ps.predict_device(torch.from_numpy(cand).cuda(), mode=1, y_mean=0.3, y_std=2.0, add_noise=True)
        wm, wv = O.mixture_of_gaussians_as_normal(want_mu * 2.0 + 0.3, want_var * 4.0 + s["noise_samples"].reshape(-1, 1))
        assert np.allclose(m1.cpu().numpy(), wm, rtol=1e-9, atol=1e-11) and np.allclose(v1.cpu().numpy(), wv, rtol=1e-8, atol=1e-11)
        e_mu, e_var = ps.predict_device(torch.empty((0, cand.shape[1]), dtype=torch.float64, device="cuda"), mode=0)
        assert e_mu.shape == (ps.num_samples, 0)
        one_mu, _ = ps.predict_device(torch.from_numpy(cand[:1]).cuda(), mode=0)
        assert np.allclose(one_mu.cpu().numpy(), want_mu[:, :1], rtol=1e-9, atol=1e-11)


def test_sampler_ragged_sizes():
    """N not a multiple of 32, a single tree, a single chain, one feature: still byte-identical to the oracle."""
    for n, dim, m, chains in ((33, 1, 1, 1), (95, 2, 3, 2), (257, 3, 5, 1)):
        want, trace_o, got, _ = replay_case(n=n, dim=dim, cat=0, m=m, chains=chains, warm=6, ns=1, sps=4, seed=n)
        assert got[0].tobytes() == want[0].tobytes()
        assert np.allclose(got[1], want[1], rtol=1e-12, atol=0)


# ------------------------------------------------------------ full-size properties (BASELINE config 4 shape)
@pytest.mark.parametrize("n,dims,m,chains", [(2000, (10, 0), 200, 4), (500, (6, 4), 100, 8)])
def test_full_size_state_properties(n, dims, m, chains):
    """BASELINE config 4's shape (N=2000, m=200, continuous; 4 of the 64 chains) and config 3's (N=500, 6 continuous
    + 4 categorical features with 5 levels, m=100; 8 of the 32 chains): after free-running sweeps the leaf-space state
    is exactly the one of the final forest (integer parts bit-exact), B^-1 is an inverse, and the running log-MLL
    equals both the GPU from-scratch evaluation and the oracle's dense Cholesky value to 1e-9."""
    X, y, bounds, ft, _ = O.synthetic_problem(n, dim=dims[0], cat_dim=dims[1], num_cat=5, m_true=50, seed=0)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    st = S.ChainState(f0, np.full(chains, 0.1), np.full(chains, 1.0), X, y, bounds, ft)
    p = B.BARKTrainParams(num_chains=chains)
    st.sweeps(p, 60, seed=11)
    r = st.read()
    assert int(r["status"].max()) == 0
    forest = st.dforest.to_numpy()
    noise, scale, run = r["noise"].cpu().numpy(), r["scale"].cpu().numpy(), r["mll"].cpu().numpy()
    scratch = B.forest_mll(forest, noise, scale, X, y, ft)
    assert np.abs(run - scratch).max() / np.abs(scratch).max() < 1e-9
    leaves_gpu = B.pass_through_forest(forest, X, ft)
    for c in (0, chains - 1):
        ex = st.export(c)
        leaves = O.pass_through_forest(forest[c], X, ft)
        assert np.array_equal(leaves_gpu[c], leaves)
        cm, P = ex["colmap"], st.p_cap
        cols = np.stack([cm[t, leaves[:, t]] for t in range(m)], axis=1)  # (n, m) leaf-space column of every point
        assert cols.min() >= 0
        Z = np.zeros((n, P))
        Z[np.arange(n)[:, None], cols] = 1
        A = (Z.T @ Z).astype(np.int32)
        assert np.array_equal(ex["A"], A)
        cc = (noise[c] + 1e-6) * m / scale[c]
        assert np.abs(ex["Binv"] @ (cc * np.eye(P) + A) - np.eye(P)).max() < 1e-7
        K = O.kernel_matrix(forest[c], X, ft, noise[c], scale[c])
        assert run[c] == pytest.approx(O.mll_cholesky(K, y), rel=1e-9)
        assert int(r["p_used"][c]) == int((forest[c]["active"] & forest[c]["is_leaf"]).sum())


def test_full_size_predict_properties():
    """Posterior predict at config-4/5 shape: both GPU kernels agree, the variance is positive and below the prior
    variance, the mixture moments equal the host formula, and a sub-sample of candidates matches the oracle."""
    n, m, chains = 2000, 200, 4
    X, y, bounds, ft, _ = O.synthetic_problem(n, dim=10, cat_dim=0, m_true=50, seed=0)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    p = B.BARKTrainParams(warmup_steps=40, num_samples=2, steps_per_sample=5, num_chains=chains)
    ns, noise, scale = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p, seed=5)
    rng = np.random.default_rng(0)
    cand = rng.random((4099, 10))  # ragged: not a multiple of the 128-candidate tile
    import torch
    ps = B.PosteriorState((ns, noise, scale), (X, y), ft, 10)
    ps_ref = B.PosteriorState((ns, noise, scale), (X, y), ft, 10, tensor_cores=False)
    cd = torch.tensor(cand, device="cuda")
    mu, var = (t.cpu().numpy() for t in ps.predict_device(cd, mode=0))
    mu2, var2 = (t.cpu().numpy() for t in ps_ref.predict_device(cd, mode=0))
    assert mu.shape == (chains * 2, 4099)
    assert np.allclose(mu, mu2, rtol=1e-12, atol=1e-12) and np.allclose(var, var2, rtol=1e-9, atol=1e-12)
    assert (var > 0).all() and (var <= (scale.reshape(-1) + 1e-9)[:, None]).all()
    mmu, mvar = (t.cpu().numpy() for t in ps.predict_device(cd, mode=1))
    hmu, hvar = B.mixture_of_gaussians_as_normal(mu, var)
    assert np.allclose(mmu, hmu, rtol=1e-12, atol=1e-12) and np.allclose(mvar, hvar, rtol=1e-9, atol=1e-12)
    sub = cand[:48]
    omu, ovar = O.forest_predict((ns[:1], noise[:1], scale[:1]), (X, y.reshape(-1, 1)), sub, ft)
    assert np.allclose(mu[:2, :48], omu, rtol=1e-9, atol=1e-9)
    assert np.allclose(var[:2, :48], ovar, rtol=1e-7, atol=1e-9)


# ------------------------------------------------------------ stochastic layer (SURVEY 8c parity protocol)
def test_posterior_statistics_match_oracle_sampler():
    """Free-running GPU chains (Philox) and free-running oracle chains (numba RNG) target the same posterior:
    chain-averaged noise, leaves per tree, and acceptance rate agree within Monte-Carlo error (BASELINE config 1
    shape: N=50, m=50, TreeFunction data; 48 chains each)."""
    n, m, chains = 50, 50, 48
    X, y, bounds, ft, _ = O.synthetic_problem(n, dim=5, cat_dim=0, m_true=50, seed=3)
    p = B.BARKTrainParams(warmup_steps=60, num_samples=8, steps_per_sample=5, num_chains=chains)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    model = (f0, np.full(chains, 0.1), np.full(chains, 1.0))
    ns_g, no_g, _, info = B.run_bark_sampler(model, (X, y), (bounds, ft), p, seed=2024, return_info=True)
    O.seed_numba(99)
    ns_o, no_o, _ = O.run_bark_sampler((f0.copy(), model[1].copy(), model[2].copy()), (X, y), bounds, ft, p)

    def per_chain(ns, no):
        leaves = (ns["active"] & ns["is_leaf"]).sum(axis=-1).mean(axis=(1, 2))  # mean leaves per tree, per chain
        depth = np.where(ns["active"] & ns["is_leaf"], ns["depth"], 0).max(axis=-1).mean(axis=(1, 2))
        return np.log(no).mean(axis=1), leaves, depth

    for a, b, name in zip(per_chain(ns_g, no_g), per_chain(ns_o, no_o), ("log noise", "leaves/tree", "max depth")):
        se = np.sqrt(a.var(ddof=1) / chains + b.var(ddof=1) / chains)
        assert abs(a.mean() - b.mean()) < 4.5 * se + 1e-3, (name, a.mean(), b.mean(), se)
    acc = info["accepted"].sum() / info["tree_proposals"].sum()
    assert 0.05 < acc < 0.9


# ------------------------------------------------------------ prior surrogate, checkpoint / resume
def test_prior_surrogate_and_checkpoint_resume(tmp_path):
    X, y, bounds, ft, _ = O.synthetic_problem(70, dim=3, cat_dim=1, num_cat=4, m_true=10, seed=4)
    y_raw = y * 1.7 - 0.3
    prior = B.BARKPriorSurrogate((bounds, ft), num_samples=4, num_trees=12, sample_seed=3)
    prior.fit(X, y_raw)
    assert prior.forest.shape == (4, 12, 100) and prior.noise.shape == (4,) and np.all(prior.scale == 1.0)
    Xc = X[:9]
    mu, sd = prior.predict(Xc)
    sc = O.Standardize(); sc.mean, sc.std = prior.scaler.mean, prior.scaler.std
    want_mu, want_sd = O.surrogate_predict(prior.model_as_tuple(), prior.train_data, Xc, ft, sc)
    assert np.allclose(mu, want_mu, rtol=1e-9, atol=1e-10) and np.allclose(sd, want_sd, rtol=1e-8, atol=1e-10)
    # posterior fit -> save -> load into a fresh surrogate -> identical predictions, and the fit resumes warm
    sur = B.BARKSurrogate((bounds, ft), warmup_steps=10, num_samples=2, steps_per_sample=3, num_trees=12, num_chains=2, seed=1)
    sur.fit(X, y_raw)
    path = tmp_path / "fit.npz"
    sur.save(path)
    sur2 = B.BARKSurrogate((bounds, ft), warmup_steps=10, num_samples=2, steps_per_sample=3, num_trees=12, num_chains=2, seed=1)
    sur2.load(path)
    assert sur2.forest.tobytes() == sur.forest.tobytes() and sur2.is_fitted
    a, b = sur.predict(Xc), sur2.predict(Xc)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    sur2.fit(X, y_raw)
    assert sur2.bark_params.warmup_steps == 0 and sur2.forest.shape == (2, 2, 12, 100)


# ------------------------------------------------------------ acquisition-model inputs
@pytest.mark.parametrize("n,dims,m", [(60, (3, 1), 12), (257, (4, 0), 40)])
def test_acquisition_inputs_match_oracle(n, dims, m):
    """K^-1, -s^2 K^-1 and s K^-1 y of the no-null kernel per posterior sample (opt_model.py:54-59,83,101) from the
    leaf-space state (Woodbury) against the oracle's dense np.linalg.inv."""
    X, y, bounds, ft, _ = O.synthetic_problem(n, dim=dims[0], cat_dim=dims[1], num_cat=4, m_true=10, seed=6)
    chains = 3
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    # few sweeps from the empty forest: some trees are still root-only, which is what "no_null" is about
    p = B.BARKTrainParams(warmup_steps=3, num_samples=2, steps_per_sample=2, num_chains=chains)
    ns, noise, scale = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p, seed=8)
    n_null = ns[..., 0]["is_leaf"].sum(axis=-1)
    assert n_null.max() > 0 and n_null.min() < m
    y_raw = (y * 3.0 + 2.0).reshape(-1, 1)
    got = B.gp_sample_inverses((ns, noise, scale), (X, y_raw), (bounds, ft))
    want = O.acquisition_inputs((ns, noise, scale), (X, y_raw), ft)
    for k in ("K_inv", "quadr_term", "lin_term", "const_term"):
        assert got[k].shape == want[k].shape, k
        err = np.abs(got[k] - want[k]).max() / np.abs(want[k]).max()
        assert err < 1e-9, (k, err)
    # an inverse is an inverse: K_inv @ K = I with K rebuilt on the GPU
    K0 = B.batched_forest_gram_matrix_no_null(ns.reshape(-1, m, 100), X, X, ft)
    K = scale.reshape(-1)[:, None, None] * K0 + (1e-6 + noise.reshape(-1))[:, None, None] * np.eye(n)
    assert np.abs(got["K_inv"] @ K - np.eye(n)).max() < 1e-8


def test_cluster_size_does_not_change_the_chain(monkeypatch):
    """1 to 16 CTAs per chain (BARK_SWEEP_CLUSTER) only change how the leaf-space linear algebra is shared out: the
    decisions are taken redundantly on identical inputs by every CTA of a chain, every 8-row unit of the DMMA product and
    every tile of the update is computed by one warp with the same arithmetic whatever the cluster size: the sampled
    forests are byte-identical and the hyper-parameter samples agree (to rounding at most)."""
    X, y, bounds, ft, _ = O.synthetic_problem(300, dim=5, cat_dim=1, num_cat=4, m_true=20, seed=12)
    chains, m = 3, 40
    p = B.BARKTrainParams(warmup_steps=25, num_samples=2, steps_per_sample=5, num_chains=chains)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    runs = {}
    for r in ("1", "2", "4", "8", "16"):
        monkeypatch.setenv("BARK_SWEEP_CLUSTER", r)
        runs[r] = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p, seed=77)
    monkeypatch.delenv("BARK_SWEEP_CLUSTER")
    for r in ("2", "4", "8", "16"):
        assert runs[r][0].tobytes() == runs["1"][0].tobytes(), f"forests differ between 1 and {r} CTAs per chain"
        assert np.allclose(runs[r][1], runs["1"][1], rtol=1e-9, atol=0)


@pytest.mark.parametrize("kb", ["1", "2", "4"])
def test_block_size_does_not_change_the_trajectory(monkeypatch, kb):
    """The speculative block size (8 proposals by default, BARK_SWEEP_KB forces the smaller instantiations that very
    wide forests fall back to) changes only how much work is batched: same oracle trajectory, byte for byte."""
    monkeypatch.setenv("BARK_SWEEP_KB", kb)
    want, trace_o, got, _ = replay_case(n=70, dim=3, cat=1, m=13, chains=2, warm=6, ns=1, sps=4, seed=31)
    monkeypatch.delenv("BARK_SWEEP_KB")
    ns_g, noise_g, scale_g, trace_g, info = got
    fin = np.isfinite(trace_o[..., 0])
    rel = np.abs(trace_g[..., 1][fin] - trace_o[..., 1][fin]) / np.maximum(np.abs(trace_o[..., 1][fin]), 1e-300)
    assert rel.max() < 1e-9, rel.max()
    assert not (trace_o[..., 2] != trace_g[..., 2]).any()
    assert ns_g.tobytes() == want[0].tobytes()


def test_sampler_wide_forest_unpaired_panels():
    """More than 512 leaf columns in use (m = 320 trees, capacity 1280: the block kernel drops to 4 proposals per block
    to fit shared memory); still the oracle's trajectory byte for byte."""
    want, trace_o, got, _ = replay_case(n=200, dim=4, cat=0, m=320, chains=2, warm=9, ns=1, sps=3, seed=21)
    ns_g, noise_g, scale_g, trace_g, info = got
    leaves = (ns_g["active"] & ns_g["is_leaf"]).sum(axis=(-1, -2))
    assert leaves.max() > 512, leaves  # the point of the test
    fin = np.isfinite(trace_o[..., 0])
    rel = np.abs(trace_g[..., 1][fin] - trace_o[..., 1][fin]) / np.maximum(np.abs(trace_o[..., 1][fin]), 1e-300)
    assert rel.max() < 1e-9, rel.max()
    assert not (trace_o[..., 2] != trace_g[..., 2]).any()
    assert ns_g.tobytes() == want[0].tobytes()


@pytest.mark.parametrize("m", [60, 130, 210, 320])
def test_predict_tensor_core_tilings(m):
    """Leaf-column extents that give 2, 3, 4 and 6 K tiles (one, two or three 256-column tiles, the last one half
    full or full) in the tcgen05 predict kernel: equal to the FP64 gather kernel and to the oracle."""
    import torch
    X, y, bounds, ft, _ = O.synthetic_problem(150, dim=4, cat_dim=1, num_cat=4, m_true=10, seed=m)
    chains = 2
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    p = B.BARKTrainParams(warmup_steps=10, num_samples=1, steps_per_sample=2, num_chains=chains)
    ns, noise, scale = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p, seed=m)
    ps = B.PosteriorState((ns, noise, scale), (X, y), ft, 5)
    assert ps.prep is not None
    ps_ref = B.PosteriorState((ns, noise, scale), (X, y), ft, 5, tensor_cores=False)
    rng = np.random.default_rng(m)
    cand = np.hstack([rng.random((301, 4)), rng.integers(0, 4, size=(301, 1)).astype(np.float64)])
    cd = torch.tensor(cand, device="cuda")
    mu, var = (t.cpu().numpy() for t in ps.predict_device(cd, mode=0))
    mu2, var2 = (t.cpu().numpy() for t in ps_ref.predict_device(cd, mode=0))
    assert np.allclose(mu, mu2, rtol=1e-12, atol=1e-12) and np.allclose(var, var2, rtol=1e-9, atol=1e-12)
    omu, ovar = O.forest_predict((ns, noise, scale), (X, y.reshape(-1, 1)), cand[:40], ft)
    assert np.allclose(mu[:, :40], omu, rtol=1e-9, atol=1e-9) and np.allclose(var[:, :40], ovar, rtol=1e-7, atol=1e-9)


def test_predict_mixture_folded_in_the_kernel():
    """Mixture mode (bark_predict_umma_mixture: samples folded in registers by the persistent predict kernel) against the
    moments formed on the host from the per-sample outputs of mode 0 with the reference's formula (bark.py:83-91), on
    more than one wave of candidate tiles, with and without the observation noise."""
    import torch
    X, y, bounds, ft, _ = O.synthetic_problem(120, dim=4, cat_dim=1, num_cat=4, m_true=10, seed=5)
    chains, m = 3, 40
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    p = B.BARKTrainParams(warmup_steps=10, num_samples=2, steps_per_sample=2, num_chains=chains)
    ns, noise, scale = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p, seed=3)
    ps = B.PosteriorState((ns, noise, scale), (X, y), ft, 5)
    assert ps.prep is not None
    rng = np.random.default_rng(0)
    n_c = 148 * 128 + 77
    cand = np.hstack([rng.random((n_c, 4)), rng.integers(0, 4, size=(n_c, 1)).astype(np.float64)])
    cd = torch.tensor(cand, device="cuda")
    mu_s, var_s = (t.cpu().numpy() for t in ps.predict_device(cd, mode=0))
    nz = np.asarray(noise, dtype=np.float64).reshape(-1)
    for y_mean, y_std, add_noise in ((0.0, 1.0, False), (1.7, 0.3, True)):
        mu, var = (t.cpu().numpy() for t in ps.predict_device(cd, mode=1, y_mean=y_mean, y_std=y_std, add_noise=add_noise))
        mj = mu_s * y_std + y_mean
        vj = var_s * (y_std * y_std) + (nz[:, None] if add_noise else 0.0)
        e = mj.mean(axis=0)
        v = (vj + mj * mj).mean(axis=0) - e * e
        assert np.allclose(mu, e, rtol=1e-13, atol=1e-13)
        assert np.allclose(var, v, rtol=1e-10, atol=1e-13)
    ps.check()


def test_tree_agreement_kernel_on_device():
    """tree_model_kernel.py:16-23 with GPU tensors in and out: equal to the oracle's forest_gram_matrix bit for bit."""
    import torch
    X, y, bounds, ft, _ = O.synthetic_problem(150, dim=3, cat_dim=1, num_cat=4, m_true=10, seed=2)
    forest = random_forests(1, 17, bounds, ft, sweeps=12, seed=4)[0]
    k = B.TreeAgreementKernel(forest, ft)
    x1 = torch.tensor(X[:97], device="cuda")
    x2 = torch.tensor(X[40:], device="cuda", dtype=torch.float64)
    got = k(x1, x2)
    assert got.is_cuda and got.shape == (97, 110)
    assert np.array_equal(got.cpu().numpy(), O.forest_gram_matrix(forest, X[:97], X[40:], ft))
    assert np.array_equal(k.forward(x1, x1).cpu().numpy(), O.forest_gram_matrix(forest, X[:97], X[:97], ft))
    assert torch.equal(k.forward(x1, x2, diag=True), torch.ones(97, dtype=torch.float64, device="cuda"))


# ------------------------------------------------------------ replay at the BASELINE shapes (configs 2, 3, 4)
def _burnt_in_start(n, dims, m, chains, burn, seed):
    """Posterior-sized forests as the common start of a full-scale replay: a short free-running GPU fit."""
    X, y, bounds, ft, _ = O.synthetic_problem(n, dim=dims[0], cat_dim=dims[1], num_cat=5, m_true=50, seed=seed)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    p = B.BARKTrainParams(warmup_steps=burn, num_samples=1, steps_per_sample=1, num_chains=chains)
    ns, noise, scale = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p,
                                          seed=seed + 7)
    return (X, y, bounds, ft), (np.ascontiguousarray(ns[:, -1]), noise[:, -1].copy(), scale[:, -1].copy())


@pytest.mark.parametrize("n,dims,m,chains,sweeps,refresh", [
    (250, (10, 0), 50, 4, 6, None),    # BASELINE config 2 shape
    (500, (6, 4), 100, 3, 3, None),    # config 3: 6 continuous + 4 categorical (bitmask splits)
    (2000, (10, 0), 200, 2, 2, None),  # config 4
    (250, (10, 0), 50, 4, 9, 2),       # the production policy: forced exact refresh (every 2 sweeps -> >= 4 per chain)
])
def test_replay_at_baseline_shapes(n, dims, m, chains, sweeps, refresh):
    """From posterior-sized forests (40 free-running sweeps) the GPU and the oracle replay the same tape: every
    proposal's log q/prior ratio, proposed log-MLL (1e-9 relative) and accept bit, and the final forests byte for
    byte -- at the sizes the bench runs, and with the forced refresh of the production path switched on."""
    (X, y, bounds, ft), start = _burnt_in_start(n, dims, m, chains, 40, seed=n)
    assert (start[0]["active"] & start[0]["is_leaf"]).sum() > 1.3 * chains * m  # the start is not the empty forest
    tape = O.make_tape(np.random.default_rng(n + 1), chains, sweeps, m)
    po = O.BARKTrainParams(warmup_steps=0, num_samples=1, steps_per_sample=sweeps, num_chains=chains)
    pg = B.BARKTrainParams(warmup_steps=0, num_samples=1, steps_per_sample=sweeps, num_chains=chains)
    trace_o = np.zeros((chains, sweeps, m + 1, 3))
    want = O.run_bark_sampler((start[0].copy(), start[1].copy(), start[2].copy()), (X, y), bounds, ft, po, tape=tape,
                              trace=trace_o)
    got = B.run_bark_sampler(start, (X, y), (bounds, ft), pg, tape=tape, return_trace=True, refresh_every=refresh)
    trace_g = got[3]
    fin = np.isfinite(trace_o[..., 0])
    assert np.array_equal(fin, np.isfinite(trace_g[..., 0]))
    tree = fin.copy(); tree[..., -1] = False
    assert np.allclose(trace_g[..., 0][tree], trace_o[..., 0][tree], rtol=0, atol=1e-12)
    assert not (trace_o[..., 2] != trace_g[..., 2]).any()
    # relative to max(|mll|, 1): along these trajectories the log-MLL passes through zero (e.g. -0.196 for one proposed
    # noise at config 2) while its terms y^T K^-1 y / 2 and log|K| / 2 are O(100); there the oracle's own LU-based value
    # carries ~1e-10 of absolute rounding, which a purely relative bound would misread as a 1e-9 relative error
    relf = np.where(fin, np.abs(trace_g[..., 1] - trace_o[..., 1]) / np.maximum(np.abs(trace_o[..., 1]), 1.0), 0.0)
    worst = np.unravel_index(np.argmax(relf), relf.shape)
    assert relf.max() < 1e-9, (relf.max(), worst, trace_g[worst], trace_o[worst], np.sort(relf.ravel())[-5:])
    assert got[0].tobytes() == want[0].tobytes()
    assert np.allclose(got[1], want[1], rtol=1e-12, atol=0)


def test_acceptance_and_heldout_statistics_match_oracle_sampler():
    """Stochastic parity at BASELINE config 2's size: free-running GPU chains (Philox) and free-running
    oracle chains (numba RNG) agree, within Monte-Carlo error, on the acceptance rates (all tree moves, change moves,
    grow/prune moves, noise moves), the share of valid proposals, and the held-out NLPD / MSE
    (src/bark/utils/metrics.py:20-39) of the posterior-predictive mixture."""
    n, n_test, m, chains = 250, 200, 50, 10
    Xa, ya, bounds, ft, _ = O.synthetic_problem(n + n_test, dim=10, cat_dim=0, m_true=50, seed=17)
    X, y, Xt, yt = np.ascontiguousarray(Xa[:n]), ya[:n].copy(), np.ascontiguousarray(Xa[n:]), ya[n:].reshape(-1)
    warm, S, sps = 50, 6, 5
    sweeps = warm + S * sps
    p = B.BARKTrainParams(warmup_steps=warm, num_samples=S, steps_per_sample=sps, num_chains=chains)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    model = (f0, np.full(chains, 0.1), np.full(chains, 1.0))
    ns_g, no_g, sc_g, tr_g = B.run_bark_sampler(model, (X, y), (bounds, ft), p, seed=4242, return_trace=True)
    tr_o = np.zeros((chains, sweeps, m + 1, 3))
    O.seed_numba(4243)
    ns_o, no_o, sc_o = O.run_bark_sampler((f0.copy(), model[1].copy(), model[2].copy()), (X, y), bounds, ft, p, trace=tr_o)

    def rates(tr):  # per chain, after warm-up
        t = tr[:, warm:, :-1]
        valid = np.isfinite(t[..., 0])
        change = valid & (t[..., 0] == 0.0)      # a valid change move has log q + log prior ratio exactly 0
        other = valid & ~change                  # grow / prune
        acc = t[..., 2] > 0
        f = lambda mask: (acc & mask).sum(axis=(1, 2)) / np.maximum(mask.sum(axis=(1, 2)), 1)
        return dict(valid=valid.mean(axis=(1, 2)), accept=acc.mean(axis=(1, 2)), accept_change=f(change),
                    accept_growprune=f(other), accept_noise=(tr[:, warm:, -1, 2] > 0).mean(axis=1))

    rg, ro = rates(tr_g), rates(tr_o)
    for k in rg:
        se = np.sqrt(rg[k].var(ddof=1) / chains + ro[k].var(ddof=1) / chains)
        assert abs(rg[k].mean() - ro[k].mean()) < 4.5 * se + 5e-3, (k, rg[k].mean(), ro[k].mean(), se)

    def heldout(ns, no, sc):  # per-chain NLPD / MSE of the mixture over that chain's samples (GPU predictor for both)
        out = []
        for c in range(chains):
            mu, var = B.forest_predict((ns[c:c + 1], no[c:c + 1], sc[c:c + 1]), (X, y), Xt, (bounds, ft))
            var = var + no[c].reshape(-1, 1)                       # predict_observed (surrogates/bark.py:85-88)
            mmu, mvar = B.mixture_of_gaussians_as_normal(mu, var)
            nl = np.mean(0.5 * np.log(2 * np.pi * mvar) + 0.5 * (yt - mmu) ** 2 / mvar)
            out.append((nl, np.mean((mmu - yt) ** 2)))
        return np.array(out)

    hg, ho = heldout(ns_g, no_g, sc_g), heldout(ns_o, no_o, sc_o)
    for k, name in enumerate(("nlpd", "mse")):
        se = np.sqrt(hg[:, k].var(ddof=1) / chains + ho[:, k].var(ddof=1) / chains)
        assert abs(hg[:, k].mean() - ho[:, k].mean()) < 4.5 * se + 1e-3, (name, hg[:, k].mean(), ho[:, k].mean(), se)


def test_running_mll_stays_within_1e9_at_low_noise():
    """Low-noise soak (ill-conditioned B = c I + Z^T Z, cond up to ~1e4): with the adaptive refresh period the
    running log-MLL carried through hundreds of rank-2 updates stays within 1e-9 (relative) of the log-MLL
    recomputed from scratch from the same forest by the point-space path (tcgen05 Gram + block LDL^T)."""
    n, m, chains = 1000, 100, 8
    X, y, bounds, ft, _ = O.synthetic_problem(n, dim=10, cat_dim=0, m_true=50, seed=23, noise_std=0.01)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    st = S.ChainState(f0, np.full(chains, 0.02), np.full(chains, 1.0), X, y, bounds, ft)
    p = B.BARKTrainParams(num_chains=chains)
    worst = 0.0
    for leg in range(6):
        st.sweeps(p, 50, seed=7, sweep_offset=50 * leg)
        r = st.read()
        assert int(r["status"].max()) == 0
        noise, scale, run = r["noise"].cpu().numpy(), r["scale"].cpu().numpy(), r["mll"].cpu().numpy()
        scratch = B.forest_mll(st.dforest.to_numpy(), noise, scale, X, y, ft)
        worst = max(worst, float((np.abs(run - scratch) / np.abs(scratch)).max()))
    assert noise.min() < 5e-3, noise  # the soak did reach the ill-conditioned regime
    assert worst < 1e-9, worst


# ------------------------------------------------------------ SURVEY 8f-2 remainder, diag=False
def test_forest_predict_full_covariance_matches_oracle():
    """`forest_predict(..., diag=False)` (src/bark/tree_kernels/tree_gps.py:107-112): the full (S, n_c, n_c) matrix
    scale - K_xX K^-1 K_Xx as the reference forms it; its diagonal is the diag=True variance."""
    X, y, bounds, ft, _ = O.synthetic_problem(80, dim=3, cat_dim=1, num_cat=4, m_true=10, seed=9)
    chains, m = 2, 14
    p = B.BARKTrainParams(warmup_steps=15, num_samples=2, steps_per_sample=3, num_chains=chains)
    f0 = np.tile(O.create_empty_forest(m), (chains, 1, 1))
    model = B.run_bark_sampler((f0, np.full(chains, 0.1), np.full(chains, 1.0)), (X, y), (bounds, ft), p, seed=3)
    cand = np.ascontiguousarray(np.vstack([X[:5], O.synthetic_problem(32, dim=3, cat_dim=1, num_cat=4, m_true=10, seed=10)[0]]))
    mu, cov = B.forest_predict(model, (X, y), cand, (bounds, ft), diag=False)
    mu_o, cov_o = O.forest_predict(model, (X, y), cand, ft, diag=False)
    assert cov.shape == (chains * 2, 37, 37)
    assert np.allclose(mu, mu_o, rtol=1e-9, atol=1e-11)
    assert np.allclose(cov, cov_o, rtol=1e-8, atol=1e-10)
    _, var = B.forest_predict(model, (X, y), cand, (bounds, ft), diag=True)
    assert np.allclose(np.diagonal(cov, axis1=1, axis2=2), var, rtol=1e-9, atol=1e-11)


def test_device_prior_sampler_law_and_structure():
    """csrc/prior.cu grows forests from the same prior as the host sampler (bark_prior_sampler.py:15-62): valid tree
    structure, rules inside the node's feasible box, and the same depth / leaf-count law as the host sampler within
    Monte-Carlo error; reproducible by seed."""
    from bark_b200 import prior
    bounds = np.array([[0.0, 1.0], [-2.0, 3.0], [0.0, 31.0], [2.0, 9.0]])
    ft = np.array([2, 2, 0, 1])
    m, S = 200, 12
    dev = B.sample_forest_prior_device(m, bounds, ft, 0.95, 2.0, S, seed=5)
    assert dev.shape == (S, m, 100) and dev.dtype == B.NODE_RECORD_DTYPE
    assert dev.tobytes() == B.sample_forest_prior_device(m, bounds, ft, 0.95, 2.0, S, seed=5).tobytes()
    assert dev.tobytes() != B.sample_forest_prior_device(m, bounds, ft, 0.95, 2.0, S, seed=6).tobytes()
    for tree in dev.reshape(-1, 100)[:400]:
        assert tree[0]["active"] and tree[0]["depth"] == 0
        for i in np.flatnonzero(tree["active"]):
            nd = tree[i]
            sub = prior.get_node_subspace(tree, int(i), bounds, ft)
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
    host = B.sample_forest_prior(m, bounds, ft, 0.95, 2.0, S, np.random.default_rng(8))

    def stats(fr):  # per sample forest: mean leaves per tree, mean max depth, share of split roots
        leaves = (fr["active"] & fr["is_leaf"]).sum(axis=-1)
        depth = np.where(fr["active"] & fr["is_leaf"], fr["depth"], 0).max(axis=-1)
        return leaves.mean(axis=1), depth.mean(axis=1), (fr[:, :, 0]["is_leaf"] == 0).mean(axis=1)

    for a, b, name in zip(stats(dev), stats(host), ("leaves/tree", "max depth", "root split")):
        se = np.sqrt(a.var(ddof=1) / S + b.var(ddof=1) / S)
        assert abs(a.mean() - b.mean()) < 4.5 * se + 1e-3, (name, a.mean(), b.mean(), se)
    sur = B.BARKPriorSurrogate((bounds, ft), num_samples=3, num_trees=10, prior_on_device=True)
    X = np.column_stack([np.random.default_rng(0).random((20, 2)), np.random.default_rng(1).integers(0, 5, 20),
                         np.random.default_rng(2).integers(2, 9, 20)]).astype(np.float64)
    mu, sd = sur.fit(X, np.arange(20.0)).predict(X[:4])
    assert mu.shape == (4, 1) and np.all(np.isfinite(mu)) and np.all(sd > 0)


def test_reference_regression_example_on_swapped_imports():
    """scripts/example_regression.py = examples/regression/regression.py:75-119 with the imports swapped: BoFire-style data
    model -> surrogate_map -> fit(experiments DataFrame) -> predict(DataFrame) -> NLPD / MSE.  The fitted surrogate must
    beat the trivial predictor (the training mean) on held-out TreeFunction data with mixed inputs."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("example_regression",
                                                  os.path.join(os.path.dirname(os.path.dirname(__file__)), "scripts",
                                                               "example_regression.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    df = ex.main(0, 150, 80, 1, dict(num_chains=4, num_trees=30, warmup_steps=60, num_samples=4, steps_per_sample=5))
    assert list(df.columns) == ["NLPD", "MSE", "Time"] and np.all(np.isfinite(df.to_numpy()))
    bench = ex.TreeFunctionBenchmark()
    y = bench.f(bench.sample(2000, 1), seed=1)["y"].to_numpy()
    assert df["MSE"].iloc[0] < 0.6 * y.var(), (df, y.var())
