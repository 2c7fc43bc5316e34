"""Multi-GPU plumbing: one process per GPU (torch.distributed), chains / candidate batches sharded with NO
data-path collective; NCCL (NVLink 5 / NVSwitch) is used only to gather posterior samples and predictive
moments afterwards.  The same code runs on gloo/CPU tensors for the host-logic tests."""
from __future__ import annotations

import numpy as np

from .forest import NODE_RECORD_DTYPE


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition of `total` units; the first `total % world` ranks get one extra."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def all_gather_ragged(local, total: int, group=None):
    """All-gather along dim 0 of per-rank blocks whose sizes follow `shard_bounds(total, r, world)`.
    `local` is a torch tensor (CUDA under NCCL, CPU under gloo); returns the (total, ...) tensor on every rank."""
    import torch
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def gather_samples(node_samples: np.ndarray, noise: np.ndarray, scale: np.ndarray, total_chains: int, device=None,
                   group=None):
    """Gather per-rank posterior samples ((C_r,S,m,L) records, (C_r,S), (C_r,S)) into the full-chain arrays."""
    import torch
    rank, world = _world(group)
    if world == 1:
        return node_samples, noise, scale
    dev = torch.device(device) if device is not None else torch.device("cpu")
    c, s, m, L = node_samples.shape
    raw = torch.from_numpy(np.ascontiguousarray(node_samples).view(np.uint8).reshape(c, -1)).to(dev)
    hyp = torch.from_numpy(np.stack([noise, scale], axis=-1).astype(np.float64)).to(dev)
    raw = all_gather_ragged(raw, total_chains, group).cpu().numpy()
    hyp = all_gather_ragged(hyp, total_chains, group).cpu().numpy()
    ns = raw.view(NODE_RECORD_DTYPE).reshape(total_chains, s, m, L)
    return ns, np.ascontiguousarray(hyp[..., 0]), np.ascontiguousarray(hyp[..., 1])


def run_bark_sampler_distributed(model, data, domain, params, *, seed: int, group=None, **kw):
    """`run_bark_sampler` over all ranks: rank r runs chains shard_bounds(C, r, world) on its GPU with Philox
    streams keyed by the GLOBAL chain index (so the result does not depend on the number of GPUs), then the
    samples are all-gathered.  Every rank returns the full (C,S,m,L), (C,S), (C,S) arrays."""
    import copy

    import torch

    from .sampler import run_bark_sampler
    rank, world = _world(group)
    forest, noise, scale = model
    total = forest.shape[0]
    if total < world:  # checked on EVERY rank before any collective, so that all ranks raise together
        raise ValueError(f"more ranks ({world}) than chains ({total})")
    lo, hi = shard_bounds(total, rank, world)
    p = copy.copy(params)
    p.num_chains = hi - lo
    if kw.get("tape") is not None:  # a replay tape covers all chains: this rank replays its own slice
        kw = dict(kw, tape=np.ascontiguousarray(np.asarray(kw["tape"])[lo:hi]))
    dev = torch.device("cuda", torch.cuda.current_device())
    ns, no, sc = run_bark_sampler((np.ascontiguousarray(forest[lo:hi]), np.reshape(noise, -1)[lo:hi],
                                   np.reshape(scale, -1)[lo:hi]), data, domain, p, seed=seed, chain_offset=lo,
                                  device=dev, **kw)
    return gather_samples(ns, no, sc, total, device=dev, group=group)


def predict_distributed(posterior_state, candidates: np.ndarray, *, mode=1, group=None, **kw):
    """Shard candidates over ranks (every rank holds all posterior samples), predict locally, all-gather the
    moments: mode 1 -> (n_c,), (n_c,); mode 0 -> (S, n_c), (S, n_c)."""
    import torch

    from .forest import _as_device_f64
    rank, world = _world(group)
    n_c = candidates.shape[0]
    lo, hi = shard_bounds(n_c, rank, world)
    dev = posterior_state.state.device
    mu, var = posterior_state.predict_device(_as_device_f64(candidates[lo:hi], dev), mode=mode, **kw)
    if mode == 0:
        mu, var = mu.t().contiguous(), var.t().contiguous()
    both = all_gather_ragged(torch.stack([mu, var], dim=-1), n_c, group)
    mu, var = both[..., 0], both[..., 1]
    if mode == 0:
        mu, var = mu.t(), var.t()
    return mu.cpu().numpy(), var.cpu().numpy()
