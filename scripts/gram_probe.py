"""Time the from-scratch Gram path (leaf columns -> one-hot tiles -> tcgen05 counts -> FP64 kernel matrix) at the bench
shape (64 forests x 2000 points x 200 trees, the posterior-sized forest of tests/golden/bench_start_c4.npz); meant to be
run plainly for the event-timed figure and once under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum` for the per-kernel split."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bark_b200 as B  # noqa: E402
from bark_b200.forest import DeviceForest, _as_device_f64, _feat_types_device, forest_slots, gram_umma_device, traverse_device  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
batch, n, m, d = 64, 2000, 200, 10
z = np.load(os.path.join(ROOT, "tests", "golden", "bench_start_c4.npz"))
forest = z["forest"].view(B.NODE_RECORD_DTYPE).reshape(m, -1)
forests = np.ascontiguousarray(np.broadcast_to(forest, (batch,) + forest.shape))
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
X = rng.random((n, d))
ft = np.ones(d, dtype=np.int64)
df = DeviceForest.from_numpy(forests, dev)
Xd, ftd = _as_device_f64(X, dev), _feat_types_device(ft, dev)
leaves = traverse_device(df, Xd, ftd)
slots = forest_slots(forests)
sc = torch.ones(batch, dtype=torch.float64, device=dev)
nz = torch.full((batch,), 0.01, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = []
for it in range(reps + 2):
    flush.fill_(it & 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, K = gram_umma_device(leaves, leaves, slots=slots, want_counts=False, scale=sc, noise=nz)
    e1.record()
    torch.cuda.synchronize()
    if it >= 2:
        ms.append(e0.elapsed_time(e1))
    del K
med = float(np.median(ms))
print(json.dumps({"what": "gram (columns + one-hot build + tcgen05 + FP64 epilogue)", "batch": batch, "n": n, "m": m, "slots": int(slots),
                  "ms": med, "out_gb": batch * n * n * 8 / 1e9, "hbm_write_gbs": batch * n * n * 8 / 1e9 / (med / 1e3)}))
