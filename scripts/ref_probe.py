"""Probe: time the staged reference sampler at BASELINE config 4 on this box's host cores (both BLAS thread modes)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_runner as R, bark_oracle as O

t = time.time(); R.warm_jit(10); print("jit s", time.time() - t, flush=True)
X, y, b, ft, _ = O.synthetic_problem(2000, dim=10, m_true=50, seed=0)
f = O.create_empty_forest(200)
out = {"cores": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
for thr in (None, 1):
    v, times, per, props = R.time_sampler_slices(f, 0.1, 1.0, X, y, b, ft, 20.0, steps=2, warmup=0, threads=thr)
    out[f"slices_threads_{thr}"] = {"proposals_per_s": v, "per_step": per, "props": props, "times": times}
    print(out, flush=True)
v, dt = R.time_full_sweep(f, 0.1, 1.0, X, y, b, ft)
out["full_sweep_default_threads"] = {"proposals_per_s": v, "seconds": dt}
print(json.dumps(out))
