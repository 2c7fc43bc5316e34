"""Diagnostic: forward block sweep (bark_mll_batched) on leaf-space-like SPD matrices vs numpy."""
import numpy as np
import torch
from bark_b200.mll import mll_batched_device

rng = np.random.default_rng(0)
for P in (20, 32, 33, 55, 64, 70, 100, 130):
    N, worst = 120, 0.0
    for rep in range(20):
        Z = np.zeros((N, P))
        col = 0
        while col < P:
            k = min(int(rng.integers(2, 5)), P - col)
            a = rng.integers(0, k, size=N)
            for i in range(k):
                Z[a == i, col + i] = 1
            col += k
        c = float(rng.uniform(0.5, 5.0))
        Bm = c * np.eye(P) + Z.T @ Z
        yv = rng.normal(size=P)
        out = mll_batched_device(torch.tensor(Bm[None], device="cuda"), torch.tensor(yv, device="cuda"))
        logdet, quad = float(out[1][0]), float(out[2][0])
        sl = np.linalg.slogdet(Bm)[1]
        qr = yv @ np.linalg.solve(Bm, yv)
        worst = max(worst, abs(logdet - sl) / abs(sl), abs(quad - qr) / abs(qr))
    print(P, "worst rel err", worst)
