"""Diagnostic (GPU box): after a bench-like burn-in, which of the running log-MLL and the from-scratch GPU log-MLL is
off, and by how much, against a host value refined in longdouble -- for the chains where the two disagree most."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bark_b200 as B  # noqa: E402
from bark_b200.sampler import ChainState  # noqa: E402
from oracle import bark_oracle as O  # noqa: E402

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
cfg = bench.CONFIGS[4]
X, y, bounds, ft, _ = bench.problem(cfg)
C = 64
st = ChainState(np.tile(B.create_empty_forest(cfg["m"]), (C, 1, 1)), np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft)
st.sweeps(B.BARKTrainParams(num_chains=C), sweeps, bench.SEED)
r = st.read()
hf = st.dforest.to_numpy()
noise, scale, run = r["noise"].cpu().numpy(), r["scale"].cpu().numpy(), r["mll"].cpu().numpy()
scratch = B.forest_mll(hf, noise, scale, X, y, ft)
rel = np.abs(run - scratch) / np.abs(scratch)
print("noise min/median/max", noise.min(), np.median(noise), noise.max(), " rel diff max", rel.max(), "median", np.median(rel))
yv = y.reshape(-1)
for c in np.argsort(-rel)[:3]:
    K = O.kernel_matrix(hf[c], X, ft, noise[c], scale[c])
    L = np.linalg.cholesky(K)
    a = np.linalg.solve(K, yv)
    for _ in range(3):
        res = (yv.astype(np.longdouble) - K.astype(np.longdouble) @ a.astype(np.longdouble)).astype(np.float64)
        a = a + np.linalg.solve(K, res)
    ref = 0.5 * (-(yv @ a) - 2 * np.log(np.diag(L)).sum())
    print(f"chain {c}: noise {noise[c]:.3g} mll {ref:.6f} | running rel {abs(run[c]-ref)/abs(ref):.2e} | scratch rel {abs(scratch[c]-ref)/abs(ref):.2e}")
