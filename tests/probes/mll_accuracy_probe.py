"""Accuracy of the running (leaf-space) log-MLL and of the from-scratch GPU MLL at bench scale, against a
float64 LAPACK Cholesky and a longdouble-refined value on the host (diagnostic; run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bark_b200 as B
from bark_b200 import synthetic
from bark_b200.sampler import ChainState
from oracle import bark_oracle as O
C, m, n = 8, 200, 2000
X, y, bounds, ft, _ = synthetic.synthetic_problem(n, dim=10, m_true=50, seed=0)
params = B.BARKTrainParams(num_chains=C)
f0 = np.tile(B.create_empty_forest(m), (C, 1, 1))
st = ChainState(f0, np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft)
for block in range(4):
    st.sweeps(params, 40, 7, sweep_offset=40 * block)
    r = st.read(); hf = st.dforest.to_numpy()
    noise, scale, run = r["noise"].cpu().numpy(), r["scale"].cpu().numpy(), r["mll"].cpu().numpy()
    scratch = B.forest_mll(hf, noise, scale, X, y, ft)
    for c in range(2):
        K = O.kernel_matrix(hf[c], X, ft, noise[c], scale[c])
        ref = O.mll_cholesky(K, y)
        # one step of iterative refinement in longdouble for the quadratic form
        L = np.linalg.cholesky(K); yv = y.reshape(-1)
        a = np.linalg.solve(K, yv); res = (yv.astype(np.longdouble) - (K.astype(np.longdouble) @ a.astype(np.longdouble))).astype(np.float64)
        a2 = a + np.linalg.solve(K, res)
        ref2 = 0.5 * (-(yv @ a2) - 2 * np.log(np.diag(L)).sum())
        print(f"sweeps {40*(block+1)} chain {c}: noise {noise[c]:.4g} ref {ref:.12f} refined {ref2:.12f} | running rel {abs(run[c]-ref2)/abs(ref2):.2e} | scratch rel {abs(scratch[c]-ref2)/abs(ref2):.2e} | lapack rel {abs(ref-ref2)/abs(ref2):.2e}")
