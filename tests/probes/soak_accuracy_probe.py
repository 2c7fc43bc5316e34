"""Long-run accuracy at bench scale (diagnostic; run on a GPU box): after many sweeps, when the sampled noise is
small and B is ill-conditioned, compare the running log-MLL and the from-scratch GPU log-MLL with a host value whose
quadratic form is iteratively refined in longdouble."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import bark_b200 as B  # noqa: E402
from bark_b200 import synthetic  # noqa: E402
from bark_b200.sampler import ChainState  # noqa: E402
from oracle import bark_oracle as O  # noqa: E402

C, m, n, sweeps = 16, 200, 2000, int(sys.argv[1]) if len(sys.argv) > 1 else 800
X, y, bounds, ft, _ = synthetic.synthetic_problem(n, dim=10, m_true=50, seed=0)
params = B.BARKTrainParams(num_chains=C)
st = ChainState(np.tile(B.create_empty_forest(m), (C, 1, 1)), np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft)
st.sweeps(params, sweeps, 42)
r = st.read()
hf = st.dforest.to_numpy()
noise, scale, run = r["noise"].cpu().numpy(), r["scale"].cpu().numpy(), r["mll"].cpu().numpy()
scratch = B.forest_mll(hf, noise, scale, X, y, ft)
yv = y.reshape(-1)
for c in np.argsort(noise)[:3]:
    K = O.kernel_matrix(hf[c], X, ft, noise[c], scale[c])
    L = np.linalg.cholesky(K)
    a = np.linalg.solve(K, yv)
    for _ in range(3):
        res = (yv.astype(np.longdouble) - K.astype(np.longdouble) @ a.astype(np.longdouble)).astype(np.float64)
        a = a + np.linalg.solve(K, res)
    ref = 0.5 * (-(yv @ a) - 2 * np.log(np.diag(L)).sum())
    print(f"chain {c}: noise {noise[c]:.3g} ref {ref:.10f} | running rel {abs(run[c]-ref)/abs(ref):.2e} | scratch rel {abs(scratch[c]-ref)/abs(ref):.2e}")
