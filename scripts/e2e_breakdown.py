"""Where the end-to-end time of run_bark_sampler goes (diagnostic; run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bark_b200 as B
from bark_b200 import synthetic
from bark_b200.sampler import ChainState
from bark_b200.forest import DeviceForest, NODE_RECORD_DTYPE

C, m, n = 64, 200, 2000
X, y, bounds, ft, _ = synthetic.synthetic_problem(n, dim=10, m_true=50, seed=0)
params = B.BARKTrainParams(num_chains=C)
f0 = np.tile(B.create_empty_forest(m), (C, 1, 1))
st = ChainState(f0, np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft)
st.sweeps(params, 100, 1)
torch.cuda.synchronize()
hf = st.dforest.to_numpy(); r = st.read(); hn, hs = r["noise"].cpu().numpy(), r["scale"].cpu().numpy()

def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); df = DeviceForest.from_numpy(hf); t1 = T()
    st2 = ChainState(hf, hn, hs, X, y, bounds, ft); t2 = T()
    st2.sweeps(params, 10, 2); t3 = T()
    raw = torch.empty(st2.dforest.n_nodes * 26, dtype=torch.uint8, device="cuda"); st2.dforest.pack_into(raw); t4 = T()
    host = raw.cpu().numpy().view(NODE_RECORD_DTYPE); t5 = T()
    print(f"rep{rep}: H2D+unpack {1e3*(t1-t0):.1f} ms | ChainState(H2D+alloc+init) {1e3*(t2-t1):.1f} | 10 sweeps {1e3*(t3-t2):.1f} | pack {1e3*(t4-t3):.1f} | D2H {1e3*(t5-t4):.1f}")
pe = B.BARKTrainParams(warmup_steps=0, num_samples=1, steps_per_sample=10, num_chains=C)
for rep in range(3):
    t0 = T(); out = B.run_bark_sampler((hf, hn, hs), (X, y), (bounds, ft), pe, seed=3); t1 = T()
    print(f"run_bark_sampler total {1e3*(t1-t0):.1f} ms")
