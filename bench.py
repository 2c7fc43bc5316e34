#!/usr/bin/env python
"""bench.py -- headline measurement of the BARK hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload fit|predict] [--config 4|2|3] [--scaling strong|weak]

Metric (BASELINE.json): MCMC proposals/sec with the full log-MLL evaluated (`--workload fit`, the default) and
posterior-predictive points/sec (`--workload predict`; also reported inside `extras` of the fit line).

fit      One STEP = one sweep of every chain = chains x (m tree MH proposals + 1 noise/scale MH proposal)
         (src/bark/fitting/bark_sampler.py:216-284).  Workload = BASELINE config 4 on synthetic TreeFunction data:
         N=2000, D=10 continuous, m=200 trees, 64 chains IN TOTAL, sharded over the GPUs (strong scaling: the split
         BASELINE names; `--scaling weak` keeps 64 chains per GPU instead).  Chains are burnt in (untimed) so the
         timed sweeps run on posterior-sized forests.  `--config 2|3` times the other BASELINE fit shapes.
predict  One STEP = mixture mean / variance over all 64 posterior samples for this rank's shard of the candidates
         (BASELINE config 5: 16 Mi candidates over 8 GPUs = 2 Mi per GPU).

Launch: N=1 plain python; N>1 one rank per GPU under torch.distributed.run (NCCL: barrier / max, and the
all-gather of samples / moments inside the end-to-end number).  Prints ONE JSON line on rank 0.
`--impl reference` times the UNMODIFIED reference (`_step_bark_sampler`'s statements issued per proposal, and
`forest_predict`) on the host cores instead -- staged under baseline/_ref by oracle/build_ref.py -- falling back to
the oracle port when the staged copy is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT_FIT, UNIT_PRED = "proposals/s", "points/s"
METRIC_FIT, METRIC_PRED = "mcmc_proposals_per_sec_full_mll", "posterior_predictive_points_per_sec"
CONFIGS = {  # BASELINE.json `configs`: n, continuous dims, categorical dims, trees, chains in total
    2: dict(n=250, d=10, cat=0, m=50, chains=16),
    3: dict(n=500, d=6, cat=4, m=100, chains=32),
    4: dict(n=2000, d=10, cat=0, m=200, chains=64),
}
SEED = 20261018


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fit", choices=["fit", "predict"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--burnin", type=int, default=120, help="untimed sweeps that bring the chains to posterior-sized forests")
    ap.add_argument("--chains", type=int, default=0, help="experiments: override the config's chain count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--predict-candidates", type=int, default=1 << 21,
                    help="candidates per GPU of the predict figure (BASELINE config 5's shard: 2 Mi)")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0, help="target CPU seconds of the bounded baseline sample")
    return ap.parse_args()


def fit_config(a):
    """The `config` object -- identical in the b200 and the reference arm."""
    c = CONFIGS[a.config]
    feats = f"D={c['d'] + c['cat']} ({c['d']} continuous" + (f" + {c['cat']} categorical, 5 levels" if c["cat"] else "") + ")"
    per = "in total, sharded over the GPUs" if a.scaling == "strong" else "per GPU"
    return {
        "workload": f"BARK MCMC fit, synthetic TreeFunction data, N={c['n']}, {feats}, m={c['m']} trees, "
                    f"{c['chains']} chains {per} (BASELINE config {a.config})",
        "n": c["n"], "features": c["d"] + c["cat"], "trees": c["m"], "chains": c["chains"], "chain_split": a.scaling,
        "step": "one sweep of every chain: chains x (m + 1) MH proposals, each with its full log-MLL",
    }


def predict_config(a):
    c = CONFIGS[4]
    return {
        "workload": f"BARK posterior-predictive mean/variance, mixture over 64 posterior samples of a config-4 fit "
                    f"(N={c['n']}, m={c['m']}), {a.predict_candidates} candidates per GPU ~ U[0,1]^10 (BASELINE config 5: "
                    f"16 Mi candidates over 8 GPUs)",
        "n": c["n"], "trees": c["m"], "posterior_samples": 64, "candidates_per_gpu": a.predict_candidates,
        "step": "mixture mean and variance of every candidate of the shard",
    }


def load_peaks():
    """Measured ceilings: HBM copy + bf16 from the driver's MEASURED_PEAKS.json, FP64 / int8 / L2 from our own
    micro-benchmark on this pool's B200 (scripts/peaks.cu -> profiles/peaks_r2.json)."""
    pk = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md: 6.65 TB/s)", "fp64_tflops": 37.16, "int8_tops": 4563.3,
          "l2_gbs": 17664.0, "own_src": "profiles/peaks_r2.json (scripts/peaks.cu, measured on this pool's B200)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            pk["hbm_gbs"] = float(json.load(f)["hbm_gbs"])
        pk["hbm_src"] = "measured (MEASURED_PEAKS.json hbm_gbs)"
    p = os.path.join(ROOT, "profiles", "peaks_r2.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        pk["fp64_tflops"], pk["int8_tops"], pk["l2_gbs"] = d["fp64_dmma_tflops"], d["int8_umma_tops"], d["l2_read_gbs"]
    return pk


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` summary."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_summary.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def problem(cfg, seed=0, cpu=False):
    """The synthetic TreeFunction data set of a config.  cpu=True (the reference arm, which must not touch the GPU product
    path) draws it with the oracle's generator: the same arrays bit for bit
    (tests/test_gpu_parity.py::test_synthetic_generator_matches_oracle)."""
    if cpu:
        from oracle import bark_oracle as O
        return O.synthetic_problem(cfg["n"], dim=cfg["d"], cat_dim=cfg["cat"], num_cat=5, m_true=50, seed=seed)
    from bark_b200 import synthetic
    return synthetic.synthetic_problem(cfg["n"], dim=cfg["d"], cat_dim=cfg["cat"], num_cat=5, m_true=50, seed=seed)


# --------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own numba code on the host cores
# --------------------------------------------------------------------------------------------------------
START_FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_start_c4.npz")


def reference_start_forest(cfg, config_id):
    """A posterior-sized start forest for the CPU arm.  Config 4: the committed fixture (chain 0 of the GPU arm's own
    burn-in, scripts/make_bench_start.py); otherwise a short CPU burn-in with the reference itself."""
    import bark_b200 as B
    if config_id == 4 and os.path.exists(START_FIXTURE):
        z = np.load(START_FIXTURE)
        return z["forest"].view(B.NODE_RECORD_DTYPE).reshape(cfg["m"], -1), float(z["noise"]), float(z["scale"]), "fixture"
    return B.create_empty_forest(cfg["m"]), 0.1, 1.0, "empty"


def cpu_fit_sample(cfg, forest, noise, scale, X, y, bounds, ft, budget_s, steps=1, warmup=0):
    """Bounded sample of the reference sampler: (proposals/s, kind, description, per-step seconds).  kind is
    "reference" when the unmodified reference modules are available (mounted or staged), else "port"."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_runner as R
    cores = os.cpu_count() or 1
    if R.available():
        R.warm_jit(cfg["d"] + cfg["cat"], cat=cfg["cat"] > 0)
        val, times, per_step, props = R.time_sampler_slices(forest, noise, scale, X, y, bounds, ft, budget_s, steps=steps,
                                                            warmup=warmup, threads=None)
        v1, _, _, p1 = R.time_sampler_slices(forest, noise, scale, X, y, bounds, ft, min(6.0, budget_s / 3), steps=1,
                                             threads=1)
        desc = (f"UNMODIFIED reference (src/bark/fitting/bark_sampler.py:233-282 statements issued per proposal: its njit "
                f"get_tree_proposal / get_leaf_vectors / low_rank_inv_update / low_rank_det_update / mll, inv + slogdet at the "
                f"noise step); 1 chain, {props} proposals ({per_step} per step) of the N={cfg['n']}, m={cfg['m']} workload from a "
                f"posterior-sized forest; BLAS threads = {cores} (all cores): {val:.1f} proposals/s; 1 BLAS thread: {v1:.1f} "
                f"proposals/s over {p1} proposals; per-chain rate (the reference runs chains serially, :147)")
        return val, "reference", desc, times
    from oracle import bark_oracle as O
    p = O.BARKTrainParams(num_chains=1)
    Xs, ys, bs, fs, _ = O.synthetic_problem(24, dim=cfg["d"], m_true=4, seed=0)
    tiny = O.CpuChain(O.create_empty_forest(3), 0.1, 1.0, (Xs, ys), bs, fs, p)
    tiny.advance(0, 3, True)
    chain = O.CpuChain(forest, noise, scale, (X, y), bounds, ft, p)
    m = forest.shape[0]
    t0 = time.perf_counter(); chain.advance(0, min(4, m), False); probe = (time.perf_counter() - t0) / min(4, m)
    per_step = max(1, min(m, int(budget_s / max(probe, 1e-9) / max(steps + warmup, 1))))
    times, props, cur = [], 0, min(4, m) % m
    for it in range(warmup + steps):
        t_end = min(m, cur + per_step)
        t0 = time.perf_counter(); k = chain.advance(cur, t_end, t_end == m); dt = time.perf_counter() - t0
        cur = 0 if t_end == m else t_end
        if it >= warmup:
            times.append(dt); props += k
    desc = (f"oracle PORT of the reference algorithm (staged reference modules not found): 1 chain, {props} proposals "
            f"({per_step} per step), N={cfg['n']}, m={cfg['m']}, BLAS threads = all cores")
    return props / sum(times), "port", desc, times


def cpu_predict_sample(X, y, bounds, ft, model, n_cand=256, n_samples=4):
    """The reference's `forest_predict` + mixture on a chunk it can hold (it forms the (S, n_c, N, m) comparison tensor
    and the n_c x n_c covariance): points/s scaled to 64 samples (cost is linear in S) -- labelled as extrapolated."""
    from oracle import ref_runner as R
    if not R.available():
        return None
    rng = np.random.default_rng(1)
    cand = rng.random((n_cand, X.shape[1]))
    sub = (np.ascontiguousarray(model[0][:n_samples]), np.asarray(model[1][:n_samples]), np.asarray(model[2][:n_samples]))
    R.time_predict_chunk((sub[0][:1], sub[1][:1], sub[2][:1]), (X, y), cand[:8], ft)  # JIT warm-up
    pts, dt = R.time_predict_chunk(sub, (X, y), cand, ft)
    S = model[1].reshape(-1).shape[0]
    return {"value": pts * n_samples / S, "unit": UNIT_PRED, "cores": os.cpu_count() or 1, "kind": "reference",
            "sample": f"UNMODIFIED forest_predict (src/bark/tree_kernels/tree_gps.py:80-113) + mixture on {n_cand} candidates x "
                      f"{n_samples} samples in {dt:.2f} s, EXTRAPOLATED linearly to {S} samples (the reference cannot hold the "
                      f"(S, n_c, N, m) tensor for more)"}


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    if a.workload == "predict":
        cfg = CONFIGS[4]
        X, y, bounds, ft, _ = problem(cfg, cpu=True)
        forest, noise, scale, src = reference_start_forest(cfg, 4)
        S = 64
        model = (np.tile(forest, (S, 1, 1)), np.full(S, noise), np.full(S, scale))
        cb = cpu_predict_sample(X, y, bounds, ft, model) or {"value": None, "unit": UNIT_PRED, "cores": os.cpu_count(),
                                                              "kind": "reference", "sample": "reference modules not staged"}
        line = {"impl": "reference", "metric": METRIC_PRED, "value": cb["value"], "unit": UNIT_PRED, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": predict_config(a), "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT_PRED, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    cfg = CONFIGS[a.config]
    X, y, bounds, ft, _ = problem(cfg, cpu=True)
    forest, noise, scale, src = reference_start_forest(cfg, a.config)
    budget = max(20.0, min(150.0, 6.0 * (a.steps + a.warmup)))
    val, kind, desc, times = cpu_fit_sample(cfg, forest, noise, scale, X, y, bounds, ft, budget, steps=a.steps, warmup=a.warmup)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC_FIT, "value": val, "unit": UNIT_FIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": fit_config(a),
        "cpu_baseline": {"value": val, "unit": UNIT_FIT, "cores": cores, "kind": kind, "sample": desc, "start_forest": src},
        "e2e": {"value": val, "unit": UNIT_FIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self, local_rank):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_fit(D, cfg, chains_local, chain_off, steps, warmup, burnin, flush_buf):
    """Build the chains, burn in, warm up, time `steps` sweeps (CUDA events per step, L2 flushed before each).
    Returns (state, ms_total on this rank, counter deltas over the timed steps, next sweep number, data)."""
    import bark_b200 as B
    from bark_b200.sampler import ChainState
    torch = D.torch
    X, y, bounds, ft, _ = problem(cfg)
    params = B.BARKTrainParams(num_chains=chains_local)
    forest0 = np.tile(B.create_empty_forest(cfg["m"]), (chains_local, 1, 1))
    st = ChainState(forest0, np.full(chains_local, 0.1), np.full(chains_local, 1.0), X, y, bounds, ft, device=D.dev)
    st.sweeps(params, burnin, SEED, chain_offset=chain_off, sweep_offset=0)
    torch.cuda.synchronize()
    sweep_no = burnin
    for _ in range(warmup):
        st.sweeps(params, 1, SEED, chain_offset=chain_off, sweep_offset=sweep_no)
        sweep_no += 1
    c_before = st.read()["counters"].cpu().numpy().astype(np.float64)
    D.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush_buf.zero_()
        ev[k][0].record()
        st.sweeps(params, 1, SEED, chain_offset=chain_off, sweep_offset=sweep_no)
        ev[k][1].record()
        sweep_no += 1
    D.barrier()
    ms_total = float(sum(e0.elapsed_time(e1) for e0, e1 in ev))
    r = st.read()
    B.sampler.raise_for_status(r["status"].cpu().numpy())
    dc = (r["counters"].cpu().numpy().astype(np.float64) - c_before).sum(axis=0)
    return st, params, ms_total, dc, sweep_no, (X, y, bounds, ft)


def kernel_rooflines(D, st, params, cfg, chain_off, sweep_no, kt, pk, ncu):
    """Per-kernel CUDA-event times (C ABI measurement call) and the bound each kernel is under."""
    n, m = cfg["n"], cfg["m"]
    cb = st.read()["counters"].cpu().numpy().astype(np.float64)
    ms_t, ms_e, ms_r = st.sweeps_timed3(params, kt, SEED, chain_offset=chain_off, sweep_offset=sweep_no)
    rr = st.read()
    d2 = (rr["counters"].cpu().numpy().astype(np.float64) - cb).sum(axis=0)
    p_used = rr["p_used"].cpu().numpy().astype(np.float64)
    wd = ((n + 31) // 32 + 3) // 4 * 4
    # sweep_block_kernel: FP64 tensor-pipe work actually issued -- the B^-1 V product runs all 8 DMMA columns
    # (2 x 8 x extent^2 flops per block) and the rank-2k update 2 x extent^2 flops per accepted proposal; L2 bytes: the lower
    # triangle is read twice by the product (row + column part, 8 B x extent^2) and read + written once by the update
    flops_sweep = 16.0 * d2[11] + 2.0 * d2[14]
    l2_sweep = 8.0 * d2[11] + 8.0 * d2[12] + 4.0 * wd * d2[13]
    t_s = ms_t / 1e3
    sweep = {"kernel": "sweep_block_kernel", "bound": "tensor", "pipe": "FP64 tensor pipe (DMMA m8n8k4)",
             "achieved": flops_sweep / t_s / 1e12, "peak": pk["fp64_tflops"], "unit": "TFLOP/s",
             "frac": flops_sweep / t_s / 1e12 / pk["fp64_tflops"], "ms_per_launch": ms_t / kt,
             "l2_gbs": l2_sweep / t_s / 1e9, "l2_frac_of_measured_l2_peak": l2_sweep / t_s / 1e9 / pk["l2_gbs"],
             "traffic": ncu.get("sweep_block_kernel", {}).get("dram_bytes_per_launch"),
             "note": "issue/latency-bound: serial MH decisions between the block-parallel passes; DMMA flops and L2 bytes are the "
                     "algorithmic work of the leaf-space formulation (DESIGN.md section 5)"}
    # hyper_eval: forward block LDL^T of P x P per chain (P^3/3 FMA = 2P^3/3 flops); refresh: full inverse (2 P^3 flops)
    flops_eval = float((2.0 / 3.0 * p_used ** 3).sum()) * kt
    flops_ref = float(np.mean(2.0 * p_used ** 3)) * d2[15]
    ev = {"kernel": "hyper_eval_kernel", "bound": "tensor", "achieved": flops_eval / (ms_e / 1e3) / 1e12, "peak": pk["fp64_tflops"],
          "unit": "TFLOP/s", "frac": flops_eval / (ms_e / 1e3) / 1e12 / pk["fp64_tflops"], "ms_per_launch": ms_e / kt,
          "traffic": ncu.get("hyper_eval_kernel", {}).get("dram_bytes_per_launch"),
          "note": "latency-bound: P/64 serial pivot blocks per chain on a 2-CTA cluster"}
    rf = {"kernel": "hyper_refresh_kernel", "bound": "tensor", "achieved": flops_ref / (ms_r / 1e3) / 1e12, "peak": pk["fp64_tflops"],
          "unit": "TFLOP/s", "frac": flops_ref / (ms_r / 1e3) / 1e12 / pk["fp64_tflops"], "ms_per_launch": ms_r / kt,
          "refreshes_per_launch": d2[15] / kt, "traffic": ncu.get("hyper_refresh_kernel", {}).get("dram_bytes_per_launch")}
    return sweep, ev, rf, (ms_t, ms_e, ms_r), d2, sweep_no + kt


def scratch_path_rooflines(D, st, cfg, data, pk, ncu):
    """The point-space ops (north_star 1-3): traversal -> int8 tcgen05 Gram (+ FP64 epilogue) -> batched FP64 block LDL^T."""
    torch = D.torch
    from bark_b200.forest import _as_device_f64, _feat_types_device, forest_slots, gram_umma_device, traverse_device
    from bark_b200.mll import mll_batched_device
    X, y, bounds, ft = data
    n, m, C = cfg["n"], cfg["m"], st.chains
    hf = st.dforest.to_numpy()
    rr = st.read()
    Xd, yd, ftd = _as_device_f64(X, D.dev), _as_device_f64(y.reshape(-1), D.dev), _feat_types_device(ft, D.dev)
    nz, sc_ = rr["noise"].to(torch.float64), rr["scale"].to(torch.float64)
    slots = forest_slots(hf)

    def run():
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); leaves = traverse_device(st.dforest, Xd, ftd); e[1].record()
        _, K = gram_umma_device(leaves, leaves, slots=slots, want_counts=False, scale=sc_, noise=nz); e[2].record()
        vals = mll_batched_device(K, yd)[0]; e[3].record()
        torch.cuda.synchronize()
        return [e[i].elapsed_time(e[i + 1]) for i in range(3)], vals
    run()
    ts = np.min([run()[0] for _ in range(3)], axis=0)
    vals = run()[1]
    t_tr, t_gr, t_ml = (float(t) for t in ts)
    d = X.shape[1]
    b_tr = C * (8.0 * n * d + 22.0 * m * hf.shape[2] + 4.0 * n * m)
    b_gr = C * 8.0 * n * n
    f_ml = C * (2.0 * n ** 3 / 3.0)
    rows = [
        {"kernel": "traverse_kernel", "bound": "hbm", "achieved": b_tr / (t_tr / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
         "frac": b_tr / (t_tr / 1e3) / 1e9 / pk["hbm_gbs"], "ms_per_launch": t_tr,
         "traffic": ncu.get("traverse_kernel", {}).get("dram_bytes_per_launch")},
        {"kernel": "gram_umma_kernel (+ leaf columns + one-hot build)", "bound": "hbm", "achieved": b_gr / (t_gr / 1e3) / 1e9,
         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": b_gr / (t_gr / 1e3) / 1e9 / pk["hbm_gbs"], "ms_per_launch": t_gr,
         "traffic": sum(ncu.get(k, {}).get("dram_bytes_per_launch") or 0.0 for k in
                        ("gram_umma_kernel", "onehot_build_kernel", "leaf_presence_kernel", "leaf_columns_kernel")) or None,
         "note": "bound = the 8 N^2-byte FP64 kernel matrix written once per forest; the K extent of the one-hot GEMM is the "
                 "forest's leaf count (columns numbered on the device), the time covers all four kernels"},
        {"kernel": "mll_batched_kernel", "bound": "tensor", "achieved": f_ml / (t_ml / 1e3) / 1e12, "peak": pk["fp64_tflops"],
         "unit": "TFLOP/s", "frac": f_ml / (t_ml / 1e3) / 1e12 / pk["fp64_tflops"], "ms_per_launch": t_ml,
         "traffic": ncu.get("mll_batched_kernel", {}).get("dram_bytes_per_launch")},
    ]
    summary = {"evals_per_s": C / (t_tr + t_gr + t_ml) * 1e3, "batch": C, "n": n, "m": m,
               "ms": {"traverse": t_tr, "gram_tcgen05_int8_incl_onehot_build": t_gr, "factorise_mll_fp64": t_ml},
               "running_vs_scratch_max_rel_diff": float((vals - rr["mll"]).abs().div(rr["mll"].abs()).max().item())}
    return rows, summary


def predict_measure(D, model, data, n_c, pk, ncu, steps=3, warmup=1, with_e2e=True):
    """Posterior-predictive mixture moments for this rank's shard: device-resident rate, end-to-end rate (pinned host
    candidates in, host moments out, NCCL all-gather of the moments at N > 1), int8 tensor-pipe roofline."""
    import bark_b200 as B
    torch = D.torch
    X, y, bounds, ft = data
    d = X.shape[1]
    ps = B.PosteriorState(model, (X, y), ft, d, device=D.dev)
    ps.check()
    S = ps.num_samples
    gen = torch.Generator(device=D.dev); gen.manual_seed(D.rank)
    cand = torch.rand((n_c, d), dtype=torch.float64, device=D.dev, generator=gen)
    for _ in range(max(1, warmup)):
        ps.predict_device(cand, mode=1)
    D.barrier()
    ms = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ps.predict_device(cand, mode=1); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms_step = D.max(float(np.mean(ms)))
    value = D.world * n_c / (ms_step / 1e3)
    m, P = ps.state.m, ps.state.p_cap
    out = {"points_per_s": value, "candidate_samples_per_s": value * S, "n_candidates_per_gpu": n_c, "posterior_samples": S,
           "ms_per_step": ms_step, "algorithmic_hbm_bytes_per_point": 8 * d + 16,
           "hbm_gbs": value / D.world * (8 * d + 16) / 1e9}
    # int8 UMMA work executed per candidate-sample: 7 digit planes x (one-hot row) x the triangular operand's tiles
    kp = getattr(ps, "k_pad", None)
    if kp:
        ops = float(ps.umma_ops_per_candidate_sample)
        tops = value / D.world * S * ops / 1e12
        out["roofline"] = {"kernel": "predict_umma_kernel", "bound": "tensor", "pipe": "int8 tcgen05 (kind::i8)", "achieved": tops,
                           "peak": pk["int8_tops"], "unit": "TOP/s", "frac": tops / pk["int8_tops"],
                           "traffic": ncu.get("predict_umma_kernel", {}).get("dram_bytes_per_launch")}
    if with_e2e:
        host = torch.empty((n_c, d), dtype=torch.float64, pin_memory=True)
        host.copy_(cand)
        out_mu = torch.empty(n_c, dtype=torch.float64, pin_memory=True)
        out_var = torch.empty(n_c, dtype=torch.float64, pin_memory=True)

        def once():
            cd = host.to(D.dev, non_blocking=True)
            mu, var = ps.predict_device(cd, mode=1)
            if D.world > 1:  # the caller wants the moments of ALL candidates: all-gather over NVLink
                g = [torch.empty_like(torch.stack([mu, var])) for _ in range(D.world)]
                D.dist.all_gather(g, torch.stack([mu, var]))
                mu, var = g[D.rank][0], g[D.rank][1]
                torch.cat([x.reshape(-1) for x in g])[:1].cpu()  # the gathered block is read on the host side
            out_mu.copy_(mu, non_blocking=True); out_var.copy_(var, non_blocking=True)
            torch.cuda.synchronize()
        once()
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            once()
        dt = D.max((time.perf_counter() - t0) / steps)
        out["e2e"] = {"value": D.world * n_c / dt, "unit": UNIT_PRED, "h2d_bytes_per_step": n_c * d * 8, "d2h_bytes_per_step": n_c * 16,
                      "what": "pinned host candidates -> H2D -> predict (mixture over all samples) -> "
                              + ("NCCL all-gather of the moments -> " if D.world > 1 else "") + "D2H of mean and variance"}
    return out


def run_b200_arm(a, local_rank):
    D = Dist(local_rank)
    torch = D.torch
    import bark_b200 as B
    from bark_b200 import distributed as BD

    pk, ncu = load_peaks(), load_ncu_traffic()
    cfg = dict(CONFIGS[a.config])
    if a.chains:
        cfg["chains"] = a.chains
    total = cfg["chains"] * (D.world if a.scaling == "weak" else 1)
    lo, hi = BD.shard_bounds(total, D.rank, D.world)
    C = hi - lo
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=D.dev)  # L2 flush between timed steps
    clocks = ClockSampler(local_rank)
    clocks.start()
    st, params, ms_total, dc, sweep_no, data = timed_fit(D, cfg, C, lo, a.steps, a.warmup, a.burnin, flush_buf)
    clk = clocks.stop()
    X, y, bounds, ft = data
    m, n = cfg["m"], cfg["n"]
    ms_total = D.max(ms_total)
    props_per_step = total * (m + 1)
    value = props_per_step * a.steps / (ms_total / 1e3)
    dc_all = np.array([D.sum(v) for v in dc])

    kt = max(4, min(a.steps, 10))
    sweep_r, eval_r, ref_r, (ms_t, ms_e, ms_r), d2, sweep_no = kernel_rooflines(D, st, params, cfg, lo, sweep_no, kt, pk, ncu)
    share = ms_t / (ms_t + ms_e + ms_r)
    sweep_r["share_of_step"] = share
    rr = st.read()
    p_used = rr["p_used"].cpu().numpy()

    # ---- end to end through the public API: host numpy in, host numpy out (at N > 1: the distributed entry point,
    # whose NCCL all-gather of the sampled forests is inside the timed call)
    host_forest = st.dforest.to_numpy()
    h_noise, h_scale = rr["noise"].cpu().numpy(), rr["scale"].cpu().numpy()
    if D.world > 1:
        full_f, full_n, full_s = BD.gather_samples(host_forest[:, None], h_noise[:, None], h_scale[:, None], total, device=D.dev)
        full_model = (np.ascontiguousarray(full_f[:, 0]), full_n[:, 0].copy(), full_s[:, 0].copy())
    else:
        full_model = (host_forest, h_noise, h_scale)
    ke = max(2, min(a.steps, 50))
    pe = B.BARKTrainParams(warmup_steps=0, num_samples=1, steps_per_sample=ke, num_chains=total)
    pw = B.BARKTrainParams(warmup_steps=0, num_samples=1, steps_per_sample=1, num_chains=total)

    def fit_call(p, seed):
        if D.world > 1:
            return BD.run_bark_sampler_distributed(full_model, (X, y), (bounds, ft), p, seed=seed)
        return B.run_bark_sampler(full_model, (X, y), (bounds, ft), p, seed=seed, device=D.dev)
    fit_call(pw, SEED + 2)  # warms the pinned-host / device caching allocators, as the warm-up steps do above
    D.barrier()
    t0 = time.perf_counter()
    ns, no, sc = fit_call(pe, SEED + 1)
    torch.cuda.synchronize()
    dt = D.max(time.perf_counter() - t0)
    h2d = (host_forest.nbytes + X.nbytes + y.nbytes + bounds.nbytes + np.asarray(ft).nbytes + h_noise.nbytes + h_scale.nbytes)
    d2h = ns.nbytes * C // total + no.nbytes + sc.nbytes
    e2e = {"value": props_per_step * ke / dt, "unit": UNIT_FIT, "h2d_bytes_per_step": h2d / ke, "d2h_bytes_per_step": d2h / ke,
           "steps": ke,
           "what": ("run_bark_sampler_distributed" if D.world > 1 else "run_bark_sampler")
                   + "(host numpy forest/X/y -> host numpy samples), warm start: H2D, state build (traversal, A, B^-1), sweeps, "
                     "packing, D2H" + (", NCCL all-gather of the sampled forests" if D.world > 1 else "")}

    # ---- secondary figures
    extras, rooflines = {}, [sweep_r, eval_r, ref_r]
    if not a.no_extras:
        try:
            rows, summ = scratch_path_rooflines(D, st, cfg, data, pk, ncu)
            rooflines += rows
            extras["full_mll_from_scratch"] = summ
        except Exception as exc:
            extras["full_mll_error"] = f"{type(exc).__name__}: {exc}"
        try:
            pm = predict_measure(D, (full_model[0][:64], full_model[1][:64], full_model[2][:64]), data, a.predict_candidates, pk, ncu)
            if "roofline" in pm:
                rooflines.append(pm.pop("roofline"))
            extras["predict"] = pm
        except Exception as exc:
            extras["predict_error"] = f"{type(exc).__name__}: {exc}"
        if a.config == 4 and D.world == 1:
            for cid in (2, 3):  # the other BASELINE fit shapes, same protocol, short
                try:
                    c2 = CONFIGS[cid]
                    s2, p2, ms2, dc2, sn2, dat2 = timed_fit(D, c2, c2["chains"], 0, 10, 3, 60, flush_buf)
                    r2 = kernel_rooflines(D, s2, p2, c2, 0, sn2, 4, pk, {})
                    extras[f"config{cid}_fit"] = {
                        "proposals_per_s": c2["chains"] * (c2["m"] + 1) * 10 / (ms2 / 1e3), "ms_per_sweep": ms2 / 10,
                        "ms_tree_sweep": r2[3][0] / 4, "ms_hyper_eval": r2[3][1] / 4, "ms_hyper_refresh": r2[3][2] / 4,
                        "tree_accept_rate": float(dc2[2] / max(dc2[0], 1)), "hyper_accept_rate": float(dc2[4] / max(dc2[3], 1)),
                        "workload": f"N={c2['n']}, m={c2['m']}, {c2['chains']} chains" + (", 6 cont + 4 cat" if c2["cat"] else "")}
                    if cid == 2:
                        _, summ2 = scratch_path_rooflines(D, s2, c2, dat2, pk, {})
                        extras["config2_batched_mll"] = summ2
                    del s2
                except Exception as exc:
                    extras[f"config{cid}_error"] = f"{type(exc).__name__}: {exc}"

    # ---- CPU baseline (rank 0, N=1 only): the reference itself on one chain from this run's chain-0 forest
    cpu_baseline = None
    if D.rank == 0 and D.world == 1 and not a.no_cpu_baseline:
        try:
            val, kind, desc, _ = cpu_fit_sample(cfg, host_forest[0], float(h_noise[0]), float(h_scale[0]), X, y, bounds, ft,
                                                a.cpu_budget_s)
            cpu_baseline = {"value": val, "unit": UNIT_FIT, "cores": os.cpu_count() or 1, "kind": kind, "sample": desc}
        except Exception as exc:  # the baseline must never take the GPU number down with it
            cpu_baseline = {"value": None, "unit": UNIT_FIT, "cores": os.cpu_count() or 1, "kind": "reference",
                            "sample": f"failed: {type(exc).__name__}: {exc}"}

    if D.rank == 0:
        wd = ((n + 31) // 32 + 3) // 4 * 4
        ws_mb = C * (float(np.mean(((p_used + 15) // 16 * 16) ** 2)) * 4 + float(np.mean(p_used)) * wd * 4) / 1e6
        cfgd = fit_config(a)
        line = {
            "metric": METRIC_FIT, "value": value, "unit": UNIT_FIT, "n_gpus": D.world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfgd,
            "run": {"chains_total": total, "chains_this_gpu": C, "burnin_sweeps": a.burnin, "proposals_per_step": props_per_step,
                    "l2": f"L2 flushed (256 MiB write) before every timed step; hot state {ws_mb:.0f} MB on this GPU",
                    "leaf_columns_per_chain_mean": float(np.mean(p_used)), "p_cap": st.p_cap,
                    "parallelism": f"chains sharded x{D.world}, no data-path collective; NCCL only gathers samples / moments"},
            "acceptance": {"tree_accept_rate": float(dc_all[2] / max(dc_all[0], 1)), "tree_valid_rate": float(dc_all[1] / max(dc_all[0], 1)),
                           "hyper_accept_rate": float(dc_all[4] / max(dc_all[3], 1))},
            "clocks": clk, "e2e": e2e, "gpu_launches": 3 * a.steps,
            "roofline": sweep_r, "rooflines": rooflines,
            "peaks": {"hbm_gbs": pk["hbm_gbs"], "hbm_source": pk["hbm_src"], "fp64_tflops": pk["fp64_tflops"],
                      "int8_tops": pk["int8_tops"], "l2_gbs": pk["l2_gbs"], "source": pk["own_src"]},
            "kernel_ms_per_step": {"sweep_block_kernel": ms_t / kt, "hyper_eval_kernel": ms_e / kt, "hyper_refresh_kernel": ms_r / kt},
            "cpu_baseline": cpu_baseline, "extras": extras,
        }
        print(json.dumps(line), flush=True)
    D.close()


def run_b200_predict(a, local_rank):
    """`--workload predict`: BASELINE config 5's per-GPU shard as the headline."""
    D = Dist(local_rank)
    import bark_b200 as B
    pk, ncu = load_peaks(), load_ncu_traffic()
    cfg = CONFIGS[4]
    # posterior samples: every rank fits the same 64 chains (deterministic Philox streams), untimed
    X, y, bounds, ft, _ = problem(cfg)
    p = B.BARKTrainParams(warmup_steps=min(a.burnin, 60), num_samples=1, steps_per_sample=1, num_chains=64)
    f0 = np.tile(B.create_empty_forest(cfg["m"]), (64, 1, 1))
    ns, no, sc = B.run_bark_sampler((f0, np.full(64, 0.1), np.full(64, 1.0)), (X, y), (bounds, ft), p, seed=SEED, device=D.dev)
    model = (np.ascontiguousarray(ns[:, -1]), no[:, -1].copy(), sc[:, -1].copy())
    clocks = ClockSampler(local_rank); clocks.start()
    pm = predict_measure(D, model, (X, y, bounds, ft), a.predict_candidates, pk, ncu, steps=a.steps, warmup=max(a.warmup, 3))
    clk = clocks.stop()
    cpu_baseline = None
    if D.rank == 0 and D.world == 1 and not a.no_cpu_baseline:
        try:
            cpu_baseline = cpu_predict_sample(X, y, bounds, ft, model)
        except Exception as exc:
            cpu_baseline = {"value": None, "unit": UNIT_PRED, "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"failed: {type(exc).__name__}: {exc}"}
    if D.rank == 0:
        roof = pm.pop("roofline", None)
        e2e = pm.pop("e2e", None)
        line = {"metric": METRIC_PRED, "value": pm["points_per_s"], "unit": UNIT_PRED, "n_gpus": D.world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": pm["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "s8 one-hot x 7 base-256 digit planes of f64 (exact int32 sums), f64 combine",
                "data": "synthetic", "config": predict_config(a), "clocks": clk, "e2e": e2e, "gpu_launches": 2 * a.steps,
                "roofline": roof, "cpu_baseline": cpu_baseline, "extras": pm,
                "l2": "inputs (candidates 8 D B each) exceed L2 at 2 Mi candidates: 168 MB per step"}
        print(json.dumps(line), flush=True)
    D.close()


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
    elif a.workload == "predict":
        run_b200_predict(a, local_rank)
    else:
        run_b200_arm(a, local_rank)


if __name__ == "__main__":
    main()
