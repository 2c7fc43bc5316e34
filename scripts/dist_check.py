"""Multi-GPU check (run under torchrun): chains sharded over ranks give byte-identical samples to a single-GPU
run with the same seed (Philox streams are keyed by the global chain index), and predict shards candidates."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bark_b200 as B
from bark_b200 import distributed as D, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, y, bounds, ft, _ = synthetic.synthetic_problem(300, dim=4, cat_dim=1, num_cat=4, m_true=10, seed=1)
C, m = 6, 20
p = B.BARKTrainParams(warmup_steps=10, num_samples=2, steps_per_sample=4, num_chains=C)
f0 = np.tile(B.create_empty_forest(m), (C, 1, 1))
model = (f0, np.full(C, 0.1), np.full(C, 1.0))
ns, no, sc = D.run_bark_sampler_distributed(model, (X, y), (bounds, ft), p, seed=99)
ok = True
if rank == 0:
    ns1, no1, sc1 = B.run_bark_sampler(model, (X, y), (bounds, ft), p, seed=99)
    ok = ns.tobytes() == ns1.tobytes() and np.array_equal(no, no1) and np.array_equal(sc, sc1)
    print("distributed fit == single-GPU fit:", ok, ns.shape)
cand = np.random.default_rng(0).uniform(size=(1001, 5)); cand[:, 4] = np.floor(cand[:, 4] * 4)
ps = B.PosteriorState((ns, no, sc), (X, y), ft, 5)
mu, var = D.predict_distributed(ps, cand, mode=1)
if rank == 0:
    mu1, var1 = ps.predict_device(torch.from_numpy(cand).cuda(), mode=1)
    ok2 = np.array_equal(mu, mu1.cpu().numpy()) and np.array_equal(var, var1.cpu().numpy())
    print("distributed predict == single-GPU predict:", ok2, mu.shape)
    ok = ok and ok2
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
