"""Soak: many sweeps at bench scale; running log-MLL vs a from-scratch GPU evaluation every `every` sweeps."""
import sys
import numpy as np
import bark_b200 as B
from bark_b200 import synthetic
from bark_b200.sampler import ChainState, raise_for_status
from bark_b200.mll import forest_mll

n, m, chains, sweeps, every = (int(a) for a in sys.argv[1:6])
X, y, bounds, ft, _ = synthetic.synthetic_problem(n, dim=10, cat_dim=0, m_true=50, seed=0)
p = B.BARKTrainParams(num_chains=chains)
f0 = np.tile(B.create_empty_forest(m), (chains, 1, 1))
cs = ChainState(f0, np.full(chains, 0.1), np.full(chains, 1.0), X, y, bounds, ft)
worst = 0.0
for s0 in range(0, sweeps, every):
    cs.sweeps(p, every, 42, sweep_offset=s0)
    r = cs.read()
    raise_for_status(r["status"].cpu().numpy())
    run = r["mll"].cpu().numpy()
    ref = forest_mll(cs.dforest.to_numpy(), r["noise"].cpu().numpy(), r["scale"].cpu().numpy(), X, y, ft)
    rel = np.abs(run - ref) / np.abs(ref)
    worst = max(worst, rel.max())
    print(s0 + every, "max rel", rel.max(), "noise min", r["noise"].min().item(), "p_used max", r["p_used"].max().item(), flush=True)
print("worst", worst)
