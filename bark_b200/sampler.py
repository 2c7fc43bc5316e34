"""`run_bark_sampler`: the reference's MCMC entry point (src/bark/fitting/bark_sampler.py:95-117) served by
the device-resident leaf-space sampler (csrc/mcmc.cu).  Same model/data/params meaning, same outputs:
`(node_samples (C,S,m,L) NODE_RECORD_DTYPE, noise_samples (C,S), scale_samples (C,S))`."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .domain import unpack_domain
from .forest import NODE_RECORD_DTYPE, DeviceForest, _as_device_f64, _feat_types_device, _ptr, _stream

TAPE_PER_TREE, TAPE_PER_HYPER = 5, 3


class BARKTrainParams:
    """Field-for-field mirror of BARKTrainParamsNumba (src/bark/fitting/bark_sampler.py:48-92); defaults from
    src/bofire_mixed/data_models/surrogates/bark.py:18-40."""

    def __init__(self, warmup_steps=50, num_samples=5, steps_per_sample=10, num_chains=1, alpha=0.95, beta=2.0,
                 proposal_weights=(0.25, 0.25, 0.5), verbose=False, use_softplus_transform=True, sample_scale=False,
                 gamma_prior_shape=1.5, gamma_prior_rate=5.0):
        self.warmup_steps = int(warmup_steps)
        self.num_samples = int(num_samples)
        self.steps_per_sample = int(steps_per_sample)
        self.num_chains = int(num_chains)
        self.alpha = float(alpha)
        self.beta = float(beta)
        self.proposal_weights = np.asarray(proposal_weights, dtype=np.float64)
        self.verbose = bool(verbose)
        self.use_softplus_transform = bool(use_softplus_transform)
        self.sample_scale = bool(sample_scale)
        self.gamma_prior_shape = float(gamma_prior_shape)
        self.gamma_prior_rate = float(gamma_prior_rate)

    def to_c(self) -> _lib.Params:
        w = np.asarray(self.proposal_weights, dtype=np.float64)
        return _lib.Params(self.alpha, self.beta, (C.c_double * 3)(*w.tolist()), self.gamma_prior_shape,
                           self.gamma_prior_rate, int(self.use_softplus_transform), int(self.sample_scale))


BARKTrainParamsNumba = BARKTrainParams  # the reference's name


def max_p_cap(n: int, d: int, m: int, node_limit: int) -> int:
    """Largest leaf-column capacity the sweep kernel's shared-memory working set allows for this problem size
    (asked of the library: `bark_mcmc_max_p_cap`); the default capacity and the overflow retry are clamped to it."""
    dims = _lib.McmcDims(1, int(n), int(d), int(m), int(node_limit), 64)
    return int(_lib.load().bark_mcmc_max_p_cap(C.byref(dims)))


def default_p_cap(m: int, node_limit: int, limit: int = 8192) -> int:
    """Leaf-column capacity: 4 leaves per tree on average (posterior forests have 1.4-2.7), multiple of 64, at most
    `limit` (see `max_p_cap`).  `run_bark_sampler` doubles it and re-runs (same seed, same trajectory) if a chain ever
    needs more."""
    cap = min(m * ((node_limit + 1) // 2), max(4 * m, 128))
    return max(64, min(limit, ((cap + 63) // 64) * 64))


def raise_for_status(status: np.ndarray):
    bits = int(np.bitwise_or.reduce(status.astype(np.int64))) if status.size else 0
    if bits & _lib.ST_TREE_OVERFLOW:
        raise OverflowError("The tree container is not large enough")  # tree_proposals.py:57-58
    if bits & _lib.ST_HYPER_MODE:
        raise NotImplementedError("You must sample the scale parameter in the log space")  # noise_scale_proposals.py:78-81
    if bits & _lib.ST_COL_OVERFLOW:
        raise _lib.BarkError("leaf-column capacity exceeded: pass a larger p_cap to run_bark_sampler")
    if bits & _lib.ST_TIMEOUT:
        raise _lib.BarkError("device pipeline wait timed out (internal error)")
    if bits & _lib.ST_NOT_SPD:
        raise np.linalg.LinAlgError("B = c I + Z^T Z lost positive definiteness")


class ChainState:
    """Device-resident state of `chains` MCMC chains (forest SoA + leaf-space workspace)."""

    def __init__(self, forest: np.ndarray, noise, scale, X, y, bounds, feat_types, p_cap=None, device=None,
                 skip_null=False):
        torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device or "cuda")
        forest = np.ascontiguousarray(forest)
        if forest.ndim != 3:
            raise ValueError("forest must have shape (chains, m, node_limit)")
        self.chains, self.m, self.L = forest.shape
        X = np.ascontiguousarray(X, dtype=np.float64)
        self.n, self.d = X.shape
        self.p_cap = int(p_cap) if p_cap else default_p_cap(self.m, self.L, max_p_cap(self.n, self.d, self.m, self.L))
        self.dims = _lib.McmcDims(self.chains, self.n, self.d, self.m, self.L, self.p_cap)
        nbytes = int(self.lib.bark_mcmc_workspace_bytes(C.byref(self.dims)))
        if nbytes == 0:
            raise ValueError(f"unsupported sampler dimensions {tuple(getattr(self.dims, f[0]) for f in self.dims._fields_)}")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.dforest = DeviceForest.from_numpy(forest, self.device)
        self.X = _as_device_f64(X, self.device)
        self.y = _as_device_f64(np.reshape(y, -1), self.device)
        self.bounds = _as_device_f64(bounds, self.device)
        self.ft = _feat_types_device(feat_types, self.device)
        noise_d = _as_device_f64(np.reshape(noise, -1), self.device)
        scale_d = _as_device_f64(np.reshape(scale, -1), self.device)
        if noise_d.numel() != self.chains or scale_d.numel() != self.chains:
            raise ValueError("noise and scale must have one entry per chain")
        # skip_null: root-only trees carry no column (the acquisition model's kernel); export-only state
        _lib.check(self.lib.bark_mcmc_init_ex(C.byref(self.dims), _ptr(self.ws), self.dforest.soa(), _ptr(self.X),
                                              _ptr(self.y), _ptr(self.bounds), _ptr(self.ft), _ptr(noise_d), _ptr(scale_d),
                                              _lib.INIT_SKIP_NULL if skip_null else 0, _stream()))

    def sweeps(self, params: BARKTrainParams, n_sweeps: int, seed: int, chain_offset=0, sweep_offset=0, tape=None,
               trace=None, refresh_every=None):
        """`refresh_every`: exact refresh of the running state every that many sweeps besides the refresh on every
        accepted noise/scale move (0 = only then, as the reference; None = the library default: 8, or 0 under a tape)."""
        cp = params.to_c()
        _lib.check(self.lib.bark_mcmc_sweeps_ex(C.byref(self.dims), _ptr(self.ws), self.dforest.soa(), C.byref(cp),
                                                int(n_sweeps), C.c_uint64(seed & (2**64 - 1)), int(chain_offset),
                                                int(sweep_offset), _ptr(tape), _ptr(trace),
                                                -1 if refresh_every is None else int(refresh_every), _stream()))

    def sweeps_timed(self, params: BARKTrainParams, n_sweeps: int, seed: int, chain_offset=0, sweep_offset=0):
        """Measurement variant: returns (ms in the tree-sweep kernel, ms in the hyper kernel), summed over sweeps."""
        cp = params.to_c()
        a, b = C.c_float(0), C.c_float(0)
        _lib.check(self.lib.bark_mcmc_sweeps_timed(C.byref(self.dims), _ptr(self.ws), self.dforest.soa(), C.byref(cp),
                                                   int(n_sweeps), C.c_uint64(seed & (2**64 - 1)), int(chain_offset),
                                                   int(sweep_offset), C.byref(a), C.byref(b), _stream()))
        return a.value, b.value

    def sweeps_timed3(self, params: BARKTrainParams, n_sweeps: int, seed: int, chain_offset=0, sweep_offset=0):
        """Measurement variant: (ms tree sweep, ms noise/scale evaluation, ms exact refresh), summed over sweeps."""
        cp = params.to_c()
        ms = (C.c_float * 3)()
        _lib.check(self.lib.bark_mcmc_sweeps_timed3(C.byref(self.dims), _ptr(self.ws), self.dforest.soa(), C.byref(cp),
                                                    int(n_sweeps), C.c_uint64(seed & (2**64 - 1)), int(chain_offset),
                                                    int(sweep_offset), ms, _stream()))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def read(self):
        """dict of per-chain device tensors: noise, scale, mll, status, counters (C,8), p_used."""
        torch = _lib.require_cuda()
        c, dev = self.chains, self.device
        out = dict(noise=torch.empty(c, dtype=torch.float64, device=dev), scale=torch.empty(c, dtype=torch.float64, device=dev),
                   mll=torch.empty(c, dtype=torch.float64, device=dev), status=torch.empty(c, dtype=torch.int32, device=dev),
                   counters=torch.empty((c, 16), dtype=torch.int64, device=dev), p_used=torch.empty(c, dtype=torch.int32, device=dev))
        _lib.check(self.lib.bark_mcmc_read(C.byref(self.dims), _ptr(self.ws), _ptr(out["noise"]), _ptr(out["scale"]),
                                           _ptr(out["mll"]), _ptr(out["status"]), _ptr(out["counters"]), _ptr(out["p_used"]),
                                           _stream()))
        return out

    def export(self, chain: int):
        """Leaf-space state of one chain as numpy arrays (verification only)."""
        torch = _lib.require_cuda()
        P, dev = self.p_cap, self.device
        wd = (((self.n + 31) // 32) + 3) // 4 * 4  # words per leaf bitset, padded to 16 bytes
        A = torch.empty((P, P), dtype=torch.int32, device=dev)
        Binv = torch.empty((P, P), dtype=torch.float64, device=dev)
        colmap = torch.empty((self.m, self.L), dtype=torch.int32, device=dev)
        bits = torch.empty((P, wd), dtype=torch.int32, device=dev)
        _lib.check(self.lib.bark_mcmc_export(C.byref(self.dims), _ptr(self.ws), int(chain), _ptr(A), _ptr(Binv),
                                             _ptr(colmap), _ptr(bits), _stream()))
        return dict(A=A.cpu().numpy(), Binv=Binv.cpu().numpy(), colmap=colmap.cpu().numpy(),
                    bits=bits.cpu().numpy().view(np.uint32))


class _ColumnOverflow(Exception):
    pass


def run_bark_sampler(model, data, domain, params: BARKTrainParams, *, seed=None, p_cap=None, tape=None,
                     return_trace=False, chain_offset=0, device=None, return_info=False, refresh_every=None):
    """See `_run_bark_sampler_once`; on leaf-column overflow the run is repeated with twice the capacity
    (deterministic: same seed / tape -> same trajectory)."""
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).astype(np.uint64) @ np.array([1, 2**32], dtype=np.uint64))
    forest = model[0]
    n_pts, n_feat = np.shape(data[0])
    limit = max_p_cap(n_pts, n_feat, forest.shape[1], forest.shape[2])
    if limit <= 0:
        raise _lib.BarkError("problem too large for the sweep kernel's shared memory (n / d / node_limit)")
    cap = int(p_cap) if p_cap else default_p_cap(forest.shape[1], forest.shape[2], limit)
    while True:
        try:
            return _run_bark_sampler_once(model, data, domain, params, seed=seed, p_cap=cap, tape=tape,
                                          return_trace=return_trace, chain_offset=chain_offset, device=device,
                                          return_info=return_info, refresh_every=refresh_every)
        except _ColumnOverflow:
            if cap >= limit:
                raise _lib.BarkError(f"leaf-column capacity exceeded at p_cap={cap}, the largest this problem size allows")
            cap = min(limit, cap * 2)


def _run_bark_sampler_once(model, data, domain, params: BARKTrainParams, *, seed=None, p_cap=None, tape=None,
                           return_trace=False, chain_offset=0, device=None, return_info=False, refresh_every=None):
    """Generate samples from the BARK posterior (src/bark/fitting/bark_sampler.py:95-117).

    model  = (forest (C,m,L) NODE_RECORD_DTYPE, noise (C,), scale (C,))
    data   = (train_x (N,D) f64, train_y (N,1) f64)
    domain = a Domain-like object (bofire or bark_b200.domain) or a `(bounds (D,2), feat_types (D,))` pair
    Returns (node_samples (C,S,m,L), noise_samples (C,S), scale_samples (C,S)) as host numpy arrays.

    Extras (keyword only): `seed` keys the Philox streams (the reference is unseeded); `tape` (C, sweeps, 5m+3)
    replays pre-drawn random numbers instead (parity tests); `return_trace` appends a (C, sweeps, m+1, 3) array of
    [log_q_prior, proposed mll, accepted]; `return_info` appends a dict with acceptance counters."""
    torch = _lib.require_cuda()
    forest, noise, scale = model
    train_x, train_y = data
    bounds, feat_types = unpack_domain(domain)
    chains = forest.shape[0]
    if chains != params.num_chains:
        raise ValueError("forest.shape[0] must equal params.num_chains")
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).astype(np.uint64) @ np.array([1, 2**32], dtype=np.uint64))
    st = ChainState(forest, noise, scale, train_x, train_y, bounds, feat_types, p_cap=p_cap, device=device)
    dev = st.device
    m, L, S = st.m, st.L, params.num_samples
    per = m * TAPE_PER_TREE + TAPE_PER_HYPER
    total_sweeps = params.warmup_steps + S * params.steps_per_sample
    tape_d = None
    if tape is not None:
        tape = np.ascontiguousarray(tape, dtype=np.float64)
        if tape.shape != (chains, total_sweeps, per):
            raise ValueError(f"tape must have shape {(chains, total_sweeps, per)}")
        tape_d = torch.from_numpy(tape).to(dev)
    traces = []

    rec_bytes = m * L * NODE_RECORD_DTYPE.itemsize
    samples = torch.empty((chains, S, rec_bytes), dtype=torch.uint8, device=dev)
    stage = torch.empty((chains, rec_bytes), dtype=torch.uint8, device=dev)
    noise_s = torch.empty((chains, S), dtype=torch.float64, device=dev)
    scale_s = torch.empty((chains, S), dtype=torch.float64, device=dev)

    def run(n_sw, s0):
        if n_sw == 0:
            return
        tp = tape_d[:, s0:s0 + n_sw].contiguous() if tape_d is not None else None
        tr = torch.zeros((chains, n_sw, m + 1, 3), dtype=torch.float64, device=dev) if return_trace else None
        st.sweeps(params, n_sw, seed, chain_offset=chain_offset, sweep_offset=s0, tape=tp, trace=tr,
                  refresh_every=refresh_every)
        if tr is not None:
            traces.append(tr)

    run(params.warmup_steps, 0)
    done = params.warmup_steps
    for k in range(S):
        run(params.steps_per_sample, done)
        done += params.steps_per_sample
        st.dforest.pack_into(stage)
        samples[:, k] = stage
        r = st.read()
        noise_s[:, k] = r["noise"]
        scale_s[:, k] = r["scale"]
    final = st.read()
    status_host = final["status"].cpu().numpy()
    if int(np.bitwise_or.reduce(status_host.astype(np.int64))) & _lib.ST_COL_OVERFLOW:
        raise _ColumnOverflow()
    raise_for_status(status_host)
    # D2H through pinned host memory (torch's caching host allocator reuses the block across calls)
    host = torch.empty(samples.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(samples, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    node_samples = host.numpy().view(NODE_RECORD_DTYPE).reshape(chains, S, m, L)
    out = [node_samples, noise_s.cpu().numpy(), scale_s.cpu().numpy()]
    if return_trace:
        out.append(torch.cat(traces, dim=1).cpu().numpy() if traces else np.zeros((chains, 0, m + 1, 3)))
    if return_info:
        cnt = final["counters"].cpu().numpy()
        out.append(dict(counters=cnt, p_used=final["p_used"].cpu().numpy(), mll=final["mll"].cpu().numpy(), seed=seed,
                        tree_proposals=cnt[:, 0], valid=cnt[:, 1], accepted=cnt[:, 2], hyper=cnt[:, 3],
                        hyper_accepted=cnt[:, 4]))
    return tuple(out)
