"""Write tests/golden/bench_start_c4.npz: chain 0 of the GPU arm's own burn-in at BASELINE config 4 (120 sweeps from the
empty forest, bench.py's seed).  `bench.py --impl reference` starts the UNMODIFIED reference sampler from this
posterior-sized forest, so that both arms time sweeps over forests of the same size.  Run on a GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bark_b200 as B  # noqa: E402
from bark_b200.sampler import ChainState  # noqa: E402

cfg = bench.CONFIGS[4]
X, y, bounds, ft, _ = bench.problem(cfg)
C = 4
st = ChainState(np.tile(B.create_empty_forest(cfg["m"]), (C, 1, 1)), np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft)
st.sweeps(B.BARKTrainParams(num_chains=C), 120, bench.SEED, chain_offset=0, sweep_offset=0)
r = st.read()
f = st.dforest.to_numpy()[0]
out = os.path.join(ROOT, "tests", "golden", "bench_start_c4.npz") if len(sys.argv) < 2 else sys.argv[1]
np.savez_compressed(out, forest=f.view(np.uint8), noise=r["noise"].cpu().numpy()[0], scale=r["scale"].cpu().numpy()[0])
print(out, "leaves", int((f["active"] & f["is_leaf"]).sum()), "noise", float(r["noise"][0]))
