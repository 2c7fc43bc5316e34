"""Predict throughput probe (diagnostic; run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bark_b200 as B
from bark_b200 import synthetic
from bark_b200.sampler import ChainState
C, m, n = 64, 200, 2000
X, y, bounds, ft, _ = synthetic.synthetic_problem(n, dim=10, m_true=50, seed=0)
params = B.BARKTrainParams(num_chains=C)
f0 = np.tile(B.create_empty_forest(m), (C, 1, 1))
st = ChainState(f0, np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft)
st.sweeps(params, 100, 1)
hf = st.dforest.to_numpy(); r = st.read(); hn, hs = r["noise"].cpu().numpy(), r["scale"].cpu().numpy()
t0 = time.perf_counter()
ps = B.PosteriorState((hf, hn, hs), (X, y), ft, 10)
torch.cuda.synchronize(); print("posterior state build", time.perf_counter() - t0)
for nc in (4096, 65536):
    cand = torch.rand((nc, 10), dtype=torch.float64, device="cuda")
    for mode in (1,):
        ps.predict_device(cand, mode=mode); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mu, var = ps.predict_device(cand, mode=mode); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"n_c={nc} mode={mode}: {ms:.2f} ms -> {nc/ms*1e3:.0f} candidates/s ({nc*C/ms*1e3:.3g} candidate-samples/s)")
