"""World-size-2 gloo tests of the multi-GPU host logic (sharding + gathers), on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bark_b200 import distributed as D
from bark_b200.forest import NODE_RECORD_DTYPE


def test_shard_bounds_partition():
    for total in (1, 7, 64, 65, 16_777_216):
        for world in (1, 2, 3, 8):
            cuts = [D.shard_bounds(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total_chains, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = D.shard_bounds(total_chains, rank, world)
        S, m, L = 2, 3, 5
        full = np.zeros((total_chains, S, m, L), dtype=NODE_RECORD_DTYPE)
        rng = np.random.default_rng(0)
        full.view(np.uint8)[:] = rng.integers(0, 256, size=full.view(np.uint8).shape, dtype=np.uint8)
        noise = rng.random((total_chains, S)); scale = rng.random((total_chains, S))
        ns, no, sc = D.gather_samples(full[lo:hi].copy(), noise[lo:hi].copy(), scale[lo:hi].copy(), total_chains)
        assert ns.tobytes() == full.tobytes() and np.array_equal(no, noise) and np.array_equal(sc, scale)
        # ragged candidate moments
        n_c = 11
        lo2, hi2 = D.shard_bounds(n_c, rank, world)
        vals = torch.arange(n_c, dtype=torch.float64)
        got = D.all_gather_ragged(torch.stack([vals[lo2:hi2], -vals[lo2:hi2]], dim=-1), n_c)
        assert torch.equal(got[:, 0], vals) and torch.equal(got[:, 1], -vals)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_chains", [4, 5])
def test_gather_world2_gloo(tmp_path, total_chains):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total_chains, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
