"""Generate `tests/golden/*.npz` by RUNNING THE UNMODIFIED REFERENCE (numba code
imported from /root/reference through `oracle/ref_shim.py`).  Run in the build
container only:  `python -m oracle.make_golden`.

The fixtures pin the oracle restatement (`oracle/bark_oracle.py`) and, through
it, the CUDA path.  Nothing here is imported at test time.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
warnings.filterwarnings("ignore")

from oracle import bark_oracle as O  # noqa: E402  (only for synthetic inputs + seeding)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def ref_params(p: O.BARKTrainParams):
    from bark.fitting.bark_sampler import BARKTrainParamsNumba
    return BARKTrainParamsNumba(
        p.warmup_steps, p.num_samples, p.steps_per_sample, p.num_chains, p.alpha, p.beta,
        np.asarray(p.proposal_weights, dtype=np.float64), p.verbose, p.use_softplus_transform,
        p.sample_scale, p.gamma_prior_shape, p.gamma_prior_rate)


def main():
    ref_shim.install()
    from numba import njit
    import bark.forest as RF
    from bark.fitting import quick_inverse as RQ
    from bark.fitting import tree_proposals as RT
    from bark.fitting import tree_traversal as RTT
    from bark.fitting import noise_scale_proposals as RN
    from bark.fitting.bark_sampler import _run_bark_sampler_multichain
    from bark.utils import bit_operations as RB

    os.makedirs(OUT, exist_ok=True)
    P = 0xFFFFFFFF

    # ---- 1. known-answer tests of SURVEY 8c (reference tests upgraded to 8 fields) ----
    kat_tree = np.array(
        [(0, 0, 0.5, 1, 2, P, 0, 1), (0, 0, 0.25, 3, 4, 0, 1, 1), (1, 0, 1.0, 0, 0, 0, 1, 1),
         (1, 0, 1.0, 0, 0, 1, 2, 1), (1, 0, 1.0, 0, 0, 1, 2, 1)], dtype=RF.NODE_RECORD_DTYPE)
    x20 = np.linspace(0, 1, 20).reshape(-1, 1)
    ft1 = np.array([2])
    kat_leaves = RF.pass_through_forest(kat_tree.reshape(1, -1), x20, ft1)
    kat_K = RF.forest_gram_matrix(kat_tree.reshape(1, -1), x20, x20, ft1)

    # tests/bark_fitting/test_quick_inverse.py:55-101 with 8-field records
    f2 = np.zeros((2, 5), dtype=RF.NODE_RECORD_DTYPE)
    f2[0, 0] = (1, 0, 0, 0, 0, P, 0, 1)
    f2[1] = kat_tree
    f2[1]["threshold"][2:] = 0.0
    new0 = f2[0].copy()
    new0[0] = (0, 0, 0.75, 1, 2, P, 0, 1)
    new0[1] = (1, 0, 0, 0, 0, 0, 1, 1)
    new0[2] = (1, 0, 0, 0, 0, 0, 1, 1)
    scale, noise = 0.5, 0.1
    K = scale * RF.forest_gram_matrix(f2, x20, x20, ft1) + noise * np.eye(20)
    K_inv = np.linalg.inv(K)
    _, ld0 = np.linalg.slogdet(K)
    amp = np.sqrt(scale / 2)
    u_cur = amp * RF.get_leaf_vectors(f2[0], x20, ft1)
    u_new = amp * RF.get_leaf_vectors(new0, x20, ft1)
    mid = RQ.low_rank_inv_update(K_inv, u_cur, subtract=True)
    ld_mid = RQ.low_rank_det_update(K_inv, u_cur, ld0, subtract=True)
    fin = RQ.low_rank_inv_update(mid, u_new)
    ld_fin = RQ.low_rank_det_update(mid, u_new, ld_mid)
    y20 = np.sin(6 * x20)
    mll_lr = RQ.mll(fin, ld_fin, y20)
    f2b = f2.copy(); f2b[0] = new0
    K2 = scale * RF.forest_gram_matrix(f2b, x20, x20, ft1) + noise * np.eye(20)
    _, ld_exact = np.linalg.slogdet(K2)
    mll_exact = RQ.mll(np.linalg.inv(K2), ld_exact, y20)

    # f32-threshold boundary + categorical bitmask walks (SURVEY 8c iii, iv)
    thr = np.float32(0.1)
    t_thr = np.zeros(3, dtype=RF.NODE_RECORD_DTYPE)
    t_thr[0] = (0, 0, thr, 1, 2, P, 0, 1); t_thr[1] = (1, 0, 0, 0, 0, 0, 1, 1); t_thr[2] = (1, 0, 0, 0, 0, 0, 1, 1)
    xb = np.array([[0.1], [float(thr)], [np.nextafter(float(thr), 1)], [np.nextafter(float(thr), 0)]])
    leaves_thr = RF.pass_through_forest(t_thr.reshape(1, -1), xb, np.array([2]))
    t_cat = t_thr.copy(); t_cat[0]["threshold"] = 6.0
    xc = np.array([[0.0], [1.0], [2.0], [3.0], [2.9]])
    leaves_cat = RF.pass_through_forest(t_cat.reshape(1, -1), xc, np.array([0]))

    np.savez(os.path.join(OUT, "kat.npz"),
             kat_tree=kat_tree.view(np.uint8), x20=x20, kat_leaves=kat_leaves, kat_K=kat_K,
             f2=f2.view(np.uint8), new0=new0.view(np.uint8), ld0=ld0, ld_fin=ld_fin, ld_exact=ld_exact,
             mll_lr=mll_lr, mll_exact=mll_exact, K_inv_fin=fin,
             t_thr=t_thr.view(np.uint8), xb=xb, leaves_thr=leaves_thr,
             t_cat=t_cat.view(np.uint8), xc=xc, leaves_cat=leaves_cat)

    # ---- 2. seeded reference sampler runs (trajectory-level pin) ----
    runs = {}
    for tag, kw, m, C, seed in [
        ("cont", dict(n=50, dim=5, cat_dim=0, m_true=50, seed=0), 50, 1, 11),       # BASELINE config 1 shape
        ("mixed", dict(n=60, dim=3, cat_dim=2, num_cat=4, m_true=10, seed=3), 12, 2, 7),
    ]:
        X, y, bounds, ft, _ = O.synthetic_problem(**kw)
        p = O.BARKTrainParams(warmup_steps=20, num_samples=2, steps_per_sample=5, num_chains=C)
        f0 = np.tile(O.create_empty_forest(m), (C, 1, 1))
        O.seed_numba(seed)
        ns, no, sc = _run_bark_sampler_multichain(
            f0.copy(), np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft, ref_params(p))
        runs[tag] = (X, y, bounds, ft, ns, no, sc)
        np.savez_compressed(os.path.join(OUT, f"sampler_{tag}.npz"), X=X, y=y, bounds=bounds, feat_types=ft,
                            m=m, chains=C, seed=seed, warmup=20, num_samples=2, steps_per_sample=5,
                            node_samples=ns.view(np.uint8), noise_samples=no, scale_samples=sc)

    # ---- 3. function-level vectors on forests taken from the reference's own samples ----
    for tag in ("cont", "mixed"):
        X, y, bounds, ft, ns, no, sc = runs[tag]
        forests = ns.reshape(-1, *ns.shape[-2:])
        rng = np.random.default_rng(5)
        # probe points: training points, exact thresholds and neighbours, random
        thr_pts = []
        for fr in forests[:2]:
            for tr in fr:
                for nd in tr:
                    if nd["active"] and not nd["is_leaf"] and ft[nd["feature_idx"]] != 0 and len(thr_pts) < 40:
                        base = X[rng.integers(X.shape[0])].copy()
                        for v in (float(nd["threshold"]), np.nextafter(float(nd["threshold"]), 2.0),
                                  np.nextafter(float(nd["threshold"]), -2.0)):
                            b2 = base.copy(); b2[nd["feature_idx"]] = v; thr_pts.append(b2)
        Xp = np.vstack([X[:20]] + ([np.array(thr_pts)] if thr_pts else []))
        leaves = np.stack([RF.pass_through_forest(f, Xp, ft) for f in forests])
        gram = RF.batched_forest_gram_matrix(forests, X, X, ft)
        gram_cross = RF.batched_forest_gram_matrix(forests, Xp, X, ft)
        gram_nonull = RF.batched_forest_gram_matrix_no_null(forests, X, X, ft)
        noise_f, scale_f = no.reshape(-1), sc.reshape(-1)
        mlls, lds = [], []
        for f, a, s in zip(forests, noise_f, scale_f):
            Kf = s * RF.forest_gram_matrix(f, X, X, ft) + (1e-6 + a) * np.eye(X.shape[0])
            _, ld = np.linalg.slogdet(Kf)
            mlls.append(RQ.mll(np.linalg.inv(Kf), ld, y)); lds.append(ld)
        # leaf vectors of first forest
        lv = [RF.get_leaf_vectors(tr, X, ft) for tr in forests[0][:6]]
        # structure queries
        term = [RTT.terminal_nodes(tr) for tr in forests[0]]
        sing = [RTT.singly_internal_nodes(tr) for tr in forests[0]]
        subs = []
        for ti, tr in enumerate(forests[0]):
            for nd in term[ti][:2]:
                subs.append((ti, int(nd), RTT.get_node_subspace(tr, nd, bounds, ft)))
        # predict: restated from tree_gps.py:80-113 with the reference's gram (gpytorch import is stubbed)
        from bark.tree_kernels.tree_gps import forest_predict, mixture_of_gaussians_as_normal
        import bark.tree_kernels.tree_gps as TG
        TG.get_feature_types_array = lambda domain: ft
        cand = Xp
        mu, var = forest_predict((ns, no, sc), (X, y), cand, None, diag=True)
        mmu, mvar = mixture_of_gaussians_as_normal(mu, var)
        np.savez_compressed(
            os.path.join(OUT, f"functions_{tag}.npz"),
            Xp=Xp, leaves=leaves, gram=gram, gram_cross=gram_cross, gram_nonull=gram_nonull,
            mll=np.array(mlls), logdet=np.array(lds),
            **{f"leafvec{i}": v for i, v in enumerate(lv)},
            term=np.array([np.pad(t, (0, 100 - len(t)), constant_values=-1) for t in term]),
            sing=np.array([np.pad(t, (0, 100 - len(t)), constant_values=-1) for t in sing]),
            sub_idx=np.array([(a, b) for a, b, _ in subs]), sub_box=np.array([c for _, _, c in subs]),
            pred_mu=mu, pred_var=var, mix_mu=mmu, mix_var=mvar)

    # ---- 4. proposal ratios under a seeded numba stream ----
    X, y, bounds, ft, ns, no, sc = runs["mixed"]
    forest = ns[0, -1]
    pr = ref_params(O.BARKTrainParams())
    O.seed_numba(123)
    out_nodes, out_lqp = [], []
    for k in range(200):
        nn, lqp = RT.get_tree_proposal(forest[k % forest.shape[0]], bounds, ft, pr)
        out_nodes.append(nn.copy()); out_lqp.append(lqp)
    O.seed_numba(321)
    nz = []
    cur = 0.1
    for k in range(50):
        (a, s), lqp = RN.get_noise_scale_proposal(cur, 1.0, pr)
        nz.append((cur, a, s, lqp)); cur = a if k % 3 else cur
    masks = []
    O.seed_numba(99)
    for k in range(100):
        masks.append(RB.sample_binary_mask(0b100101))
    np.savez_compressed(os.path.join(OUT, "proposals_mixed.npz"),
                        forest=forest.view(np.uint8), prop_nodes=np.stack(out_nodes).view(np.uint8),
                        prop_lqp=np.array(out_lqp), noise_walk=np.array(nz), masks=np.array(masks))
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
