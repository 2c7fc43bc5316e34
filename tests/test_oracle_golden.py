"""The oracle restatement (oracle/bark_oracle.py) against fixtures produced by
RUNNING THE REFERENCE (oracle/make_golden.py).  CPU only; runs everywhere."""
import os

import numpy as np
import pytest

from oracle import bark_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def rec(a):
    return np.ascontiguousarray(a).view(O.NODE_RECORD_DTYPE).reshape(a.shape[:-1] + (a.shape[-1] // 26,))


def load(name):
    return np.load(os.path.join(G, name))


def test_record_layout():
    # packed 26-byte AoS records, offsets of src/bark/forest.py:8-19
    d = O.NODE_RECORD_DTYPE
    assert d.itemsize == 26
    assert [d.fields[n][1] for n in d.names] == [0, 1, 5, 9, 13, 17, 21, 25]
    f = O.create_empty_forest(3)
    assert f.shape == (3, 100) and f[0, 0]["parent"] == 0xFFFFFFFF and f[0, 0]["is_leaf"] == 1


def test_kat_five_node_tree():
    k = load("kat.npz")
    tree = rec(k["kat_tree"])
    ft = np.array([2])
    leaves = O.pass_through_forest(tree.reshape(1, -1), k["x20"], ft)
    assert np.array_equal(leaves, k["kat_leaves"])
    assert leaves[:, 0].tolist() == [3] * 5 + [4] * 5 + [2] * 10
    K = O.forest_gram_matrix(tree.reshape(1, -1), k["x20"], k["x20"], ft)
    assert np.array_equal(K, k["kat_K"]) and K.sum() == 150


def test_kat_threshold_and_bitmask_edges():
    k = load("kat.npz")
    assert O.pass_through_forest(rec(k["t_thr"]).reshape(1, -1), k["xb"], np.array([2]))[:, 0].tolist() == [1, 1, 2, 1]
    assert np.array_equal(O.pass_through_forest(rec(k["t_thr"]).reshape(1, -1), k["xb"], np.array([2])), k["leaves_thr"])
    lc = O.pass_through_forest(rec(k["t_cat"]).reshape(1, -1), k["xc"], np.array([0]))
    assert lc[:, 0].tolist() == [2, 1, 1, 2, 1] and np.array_equal(lc, k["leaves_cat"])


def test_kat_low_rank_update_with_forest():
    # reference tests/bark_fitting/test_quick_inverse.py:55-101 (8-field records)
    k = load("kat.npz")
    f2, new0 = rec(k["f2"]), rec(k["new0"])
    x, ft = k["x20"], np.array([2])
    scale, noise = 0.5, 0.1
    K = scale * O.forest_gram_matrix(f2, x, x, ft) + noise * np.eye(20)
    Kinv = np.linalg.inv(K)
    _, ld = np.linalg.slogdet(K)
    assert ld == pytest.approx(float(k["ld0"]), rel=1e-13)
    amp = np.sqrt(scale / 2)
    uc = amp * O.get_leaf_vectors(f2[0], x, ft)
    un = amp * O.get_leaf_vectors(new0, x, ft)
    mid = O.low_rank_inv_update(Kinv, uc, True)
    ldm = O.low_rank_det_update(Kinv, uc, ld, True)
    fin = O.low_rank_inv_update(mid, un, False)
    ldf = O.low_rank_det_update(mid, un, ldm, False)
    assert ldf == pytest.approx(float(k["ld_fin"]), rel=1e-13)
    assert ldf == pytest.approx(float(k["ld_exact"]), rel=1e-12)
    assert np.allclose(fin, k["K_inv_fin"], rtol=0, atol=1e-12)
    y = np.sin(6 * x)
    assert O.mll(fin, ldf, y) == pytest.approx(float(k["mll_lr"]), rel=1e-13)
    assert float(k["mll_lr"]) == pytest.approx(5.920170615417254, rel=1e-12)


@pytest.mark.parametrize("seed", [42, 43, 44])
def test_low_rank_updates_random(seed):
    # reference tests/bark_fitting/test_quick_inverse.py:13-52
    n, b = {42: (5, 2), 43: (4, 2), 44: (6, 3)}[seed]
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    U = rng.standard_normal((n, b)) * 0.1
    Ainv = np.linalg.inv(A)
    _, ld = np.linalg.slogdet(A)
    for sub, B in ((False, A + U @ U.T), (True, A - U @ U.T)):
        assert np.isclose(O.low_rank_inv_update(Ainv, U, sub), np.linalg.inv(B)).all()
        assert np.isclose(O.low_rank_det_update(Ainv, U, ld, sub), np.linalg.slogdet(B)[1])


def test_sample_binary_mask_support():
    # reference tests/test_bit_operations.py:4-17
    empty = np.empty(0)
    for _ in range(100):
        x = O.sample_binary_mask(0b100101, empty, 0)
        assert x != 0 and x != 0b100101 and (x & 0b100101) == x
        assert O.sample_binary_mask(0b00100, empty, 0) == 0
    rng = np.random.default_rng(0)
    for u in rng.random(200):
        x = O.sample_binary_mask(0b100101, np.array([u]), 0)
        assert x != 0 and x != 0b100101 and (x & 0b100101) == x


def test_sample_binary_mask_stream():
    p = load("proposals_mixed.npz")
    O.seed_numba(99)
    got = [O.sample_binary_mask(0b100101, np.empty(0), 0) for _ in range(100)]
    assert got == p["masks"].tolist()


def test_integer_split_rule():
    # reference tests/bark_fitting/test_tree_proposals.py:7-20
    ft = np.array([1])
    empty = np.empty(0)
    for _ in range(100):
        f, thr = O.sample_splitting_rule(np.array([[0.0, 10.0]]), ft, empty, 0, 0)
        assert 0 <= thr < 10
        f, thr = O.sample_splitting_rule(np.array([[5.0, 5.0]]), ft, empty, 0, 0)
        assert thr == 5


@pytest.mark.parametrize("tag", ["cont", "mixed"])
def test_function_vectors(tag):
    s = load(f"sampler_{tag}.npz")
    g = load(f"functions_{tag}.npz")
    X, y, bounds, ft = s["X"], s["y"], s["bounds"], s["feat_types"]
    ns = rec(s["node_samples"])
    forests = ns.reshape(-1, *ns.shape[-2:])
    leaves = np.stack([O.pass_through_forest(f, g["Xp"], ft) for f in forests])
    assert leaves.dtype == np.uint32 and np.array_equal(leaves, g["leaves"])
    assert np.array_equal(O.batched_forest_gram_matrix(forests, X, X, ft), g["gram"])  # bit-exact K0
    assert np.array_equal(O.batched_forest_gram_matrix(forests, g["Xp"], X, ft), g["gram_cross"])
    assert np.allclose(O.batched_forest_gram_matrix_no_null(forests, X, X, ft), g["gram_nonull"], rtol=1e-15, atol=0)
    noise, scale = s["noise_samples"].reshape(-1), s["scale_samples"].reshape(-1)
    for i, f in enumerate(forests):
        K = O.kernel_matrix(f, X, ft, noise[i], scale[i])
        assert O.mll_from_kernel(K, y) == pytest.approx(float(g["mll"][i]), rel=1e-13)
        assert O.mll_cholesky(K, y) == pytest.approx(float(g["mll"][i]), rel=1e-11)
    for i in range(6):
        assert np.array_equal(O.get_leaf_vectors(forests[0][i], X, ft), g[f"leafvec{i}"])
    for t, tr in enumerate(forests[0]):
        a = O.terminal_nodes(tr)
        b = O.singly_internal_nodes(tr)
        assert a.tolist() == [v for v in g["term"][t] if v >= 0]
        assert b.tolist() == [v for v in g["sing"][t] if v >= 0]
    for (t, nd), box in zip(g["sub_idx"], g["sub_box"]):
        assert np.array_equal(O.get_node_subspace(forests[0][t], nd, bounds, ft), box)
    mu, var = O.forest_predict((ns, s["noise_samples"], s["scale_samples"]), (X, y), g["Xp"], ft)
    assert np.allclose(mu, g["pred_mu"], rtol=1e-10, atol=1e-12)
    assert np.allclose(var, g["pred_var"], rtol=1e-9, atol=1e-12)
    mm, mv = O.mixture_of_gaussians_as_normal(mu, var)
    assert np.allclose(mm, g["mix_mu"], rtol=1e-10, atol=1e-12) and np.allclose(mv, g["mix_var"], rtol=1e-9, atol=1e-12)


def test_proposal_stream():
    """Same numba seed -> same proposals and log ratios as the reference."""
    s = load("sampler_mixed.npz")
    p = load("proposals_mixed.npz")
    forest = rec(p["forest"])
    want_nodes = rec(p["prop_nodes"])
    cdf = np.cumsum(np.array([0.25, 0.25, 0.5]))
    O.seed_numba(123)
    for k in range(200):
        nn, lqp, st = O.get_tree_proposal(forest[k % forest.shape[0]], s["bounds"], s["feat_types"],
                                          0.95, 2.0, cdf, np.empty(0), 0)
        assert st == 0 and nn.tobytes() == want_nodes[k].tobytes()
        if np.isfinite(p["prop_lqp"][k]):
            assert lqp == pytest.approx(float(p["prop_lqp"][k]), rel=1e-14, abs=1e-15)
        else:
            assert lqp == -np.inf
    O.seed_numba(321)
    for cur, a, sc, lqp in p["noise_walk"]:
        na, nsc, got, st = O.get_noise_scale_proposal(cur, 1.0, True, False, 1.5, 5.0, np.empty(0), 0)
        assert na == pytest.approx(a, rel=1e-14) and nsc == sc and got == pytest.approx(lqp, rel=1e-12, abs=1e-13)


@pytest.mark.parametrize("tag", ["cont", "mixed"])
def test_sampler_trajectory(tag):
    """Whole seeded MCMC run: byte-identical posterior samples to the reference."""
    s = load(f"sampler_{tag}.npz")
    C, m = int(s["chains"]), int(s["m"])
    p = O.BARKTrainParams(warmup_steps=int(s["warmup"]), num_samples=int(s["num_samples"]),
                          steps_per_sample=int(s["steps_per_sample"]), num_chains=C)
    f0 = np.tile(O.create_empty_forest(m), (C, 1, 1))
    O.seed_numba(int(s["seed"]))
    ns, no, sc = O.run_bark_sampler((f0, np.full(C, 0.1), np.full(C, 1.0)), (s["X"], s["y"]),
                                    s["bounds"], s["feat_types"], p)
    assert ns.tobytes() == rec(s["node_samples"]).tobytes()
    assert np.array_equal(no, s["noise_samples"]) and np.array_equal(sc, s["scale_samples"])


def test_tape_mode_runs_and_traces():
    X, y, bounds, ft, _ = O.synthetic_problem(40, dim=3, cat_dim=1, num_cat=4, m_true=8, seed=1)
    C, m, sweeps = 2, 6, 8
    p = O.BARKTrainParams(warmup_steps=4, num_samples=2, steps_per_sample=2, num_chains=C)
    tape = O.make_tape(np.random.default_rng(0), C, sweeps, m)
    trace = np.zeros((C, sweeps, m + 1, 3))
    f0 = np.tile(O.create_empty_forest(m), (C, 1, 1))
    a = O.run_bark_sampler((f0.copy(), np.full(C, 0.1), np.full(C, 1.0)), (X, y), bounds, ft, p, tape=tape, trace=trace)
    b = O.run_bark_sampler((f0.copy(), np.full(C, 0.1), np.full(C, 1.0)), (X, y), bounds, ft, p, tape=tape)
    assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])  # deterministic given the tape
    assert trace[..., 2].sum() > 0 and np.isfinite(trace[..., 1]).all()
    # running MLL of the last sweep == refactorised MLL of the final state (Woodbury drift is tiny)
    K = O.kernel_matrix(a[0][0, -1], X, ft, a[1][0, -1], a[2][0, -1])
    acc = trace[0, -1, :, 2] > 0
    last_mll = trace[0, -1, acc, 1][-1] if acc.any() else None
    if last_mll is not None:
        assert O.mll_from_kernel(K, y) == pytest.approx(last_mll, rel=1e-9)
