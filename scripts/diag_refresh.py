"""Diagnostic: accuracy of a freshly refreshed B^-1 (cluster FULL sweep) and of chain init (solo FULL sweep)."""
import numpy as np
import bark_b200 as B
from bark_b200 import synthetic
from bark_b200.sampler import ChainState

C, m, n = 8, 200, 2000
X, y, bounds, ft, _ = synthetic.synthetic_problem(n, dim=10, m_true=50, seed=0)
params = B.BARKTrainParams(num_chains=C)
f0 = np.tile(B.create_empty_forest(m), (C, 1, 1))
st = ChainState(f0, np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft)
for total in (40, 80, 120, 160):
    st.sweeps(params, 40, 7, sweep_offset=total - 40)
    r = st.read()
    noise, scale = r["noise"].cpu().numpy(), r["scale"].cpu().numpy()
    c = int((-total) % 8)  # chain refreshed at the end of the last sweep: (sweep + 1 + chain) % 8 == 0
    ex = st.export(c)
    P = ex["A"].shape[0]
    cc = (noise[c] + 1e-6) * m / scale[c]
    Bm = cc * np.eye(P) + ex["A"].astype(np.float64)
    res = np.abs(ex["Binv"] @ Bm - np.eye(P)).max()
    ref = np.linalg.inv(Bm)
    print(f"sweeps {total} chain {c}: c={cc:.4g} |Binv B - I| {res:.2e}  |Binv - inv| rel {np.abs(ex['Binv'] - ref).max() / np.abs(ref).max():.2e}  lapack resid {np.abs(ref @ Bm - np.eye(P)).max():.2e}")
    # a re-initialised state from the same forest: solo FULL sweep
    st2 = ChainState(st.dforest.to_numpy()[c:c + 1], noise[c:c + 1], scale[c:c + 1], X, y, bounds, ft)
    ex2 = st2.export(0)
    Bm2 = cc * np.eye(P) + ex2["A"].astype(np.float64)
    print(f"    init (solo): |Binv B - I| {np.abs(ex2['Binv'] @ Bm2 - np.eye(P)).max():.2e}")
