"""Diagnostic: after every sweep compare the running log-MLL of each chain with a from-scratch GPU evaluation."""
import sys
import numpy as np
import bark_b200 as B
from bark_b200.sampler import ChainState
from bark_b200.mll import forest_mll
from bark_b200 import synthetic

n, m, chains, sweeps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
X, y, bounds, ft, _ = synthetic.synthetic_problem(n, dim=4, cat_dim=0, m_true=10, seed=7)
p = B.BARKTrainParams(num_chains=chains)
f0 = np.tile(B.create_empty_forest(m), (chains, 1, 1))
cs = ChainState(f0, np.full(chains, 0.1), np.full(chains, 1.0), X, y, bounds, ft)
prev_acc = None
for s in range(sweeps):
    cs.sweeps(p, 1, 42, sweep_offset=s)
    r = cs.read()
    run = r["mll"].cpu().numpy()
    forest = cs.dforest.to_numpy()
    ref = forest_mll(forest, r["noise"].cpu().numpy(), r["scale"].cpu().numpy(), X, y, ft)
    rel = np.abs(run - ref) / np.abs(ref)
    acc = r["counters"].cpu().numpy()[:, 4]
    flag = "" if rel.max() < 1e-9 else "  <-- MISMATCH"
    if flag or s % 10 == 0:
        print(s, "max rel", rel.max(), "chain", int(rel.argmax()), "p_used", r["p_used"].cpu().numpy().max(),
              "hyper acc", None if prev_acc is None else (acc - prev_acc).tolist(), "status", r["status"].cpu().numpy().max(), flag)
    prev_acc = acc
    if flag and rel.max() > 1e-6:
        break
