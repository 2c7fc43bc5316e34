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


def _worker_sampler_plumbing(rank, world, port, out_dir):
    """run_bark_sampler_distributed's host logic on gloo with the GPU fit replaced by a recording stub: every rank gets
    its contiguous chain block, the GLOBAL chain offset for the Philox streams, ITS slice of a replay tape, and the
    gathered result is the full-chain array in chain order on every rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bark_b200.sampler as S
        total, S_, m, L = 5, 2, 3, 4
        seen = {}

        def fake_fit(model, data, domain, params, *, seed, chain_offset, device, tape=None, **kw):
            forest, noise, scale = model
            c = forest.shape[0]
            seen.update(chains=c, offset=chain_offset, params_chains=params.num_chains, tape=None if tape is None else tape.copy())
            ns = np.zeros((c, S_, m, L), dtype=NODE_RECORD_DTYPE)
            ns["depth"] = (chain_offset + np.arange(c))[:, None, None, None]  # tag every record with its global chain
            return ns, np.tile(noise[:, None], (1, S_)), np.tile(scale[:, None], (1, S_))

        S.run_bark_sampler, orig = fake_fit, S.run_bark_sampler
        torch.cuda.current_device = lambda: 0  # the distributed entry only needs a device index for its tensors
        real_gather = D.gather_samples
        D.gather_samples = lambda ns, no, sc, tot, device=None, group=None: real_gather(ns, no, sc, tot, device="cpu", group=group)
        try:
            forest = np.zeros((total, m, L), dtype=NODE_RECORD_DTYPE)
            noise, scale = np.arange(total) + 0.5, np.arange(total) + 10.0
            tape = np.arange(total * 7 * 4, dtype=np.float64).reshape(total, 7, 4)
            params = S.BARKTrainParams(num_chains=total)
            ns, no, sc = D.run_bark_sampler_distributed((forest, noise, scale), (None, None), None, params, seed=3, tape=tape)
        finally:
            S.run_bark_sampler = orig
            D.gather_samples = real_gather
        lo, hi = D.shard_bounds(total, rank, world)
        assert seen["chains"] == hi - lo == seen["params_chains"] and seen["offset"] == lo
        assert np.array_equal(seen["tape"], tape[lo:hi])
        assert params.num_chains == total  # the caller's params object is not modified
        assert ns.shape == (total, S_, m, L) and np.array_equal(ns["depth"][:, 0, 0, 0], np.arange(total))
        assert np.array_equal(no[:, 0], noise) and np.array_equal(sc[:, 1], scale)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_distributed_sampler_plumbing_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker_sampler_plumbing, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
