#!/usr/bin/env python
"""bench.py -- headline measurement of the BARK hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): MCMC proposals/sec with the full log-MLL evaluated.  One STEP = one sweep of every
resident chain = chains x (m tree MH proposals + 1 noise/scale MH proposal) (bark_sampler.py:216-284).
Workload: BASELINE config 4 on synthetic TreeFunction data -- N=2000 training points, D=10 continuous features,
m=200 trees, 64 chains per GPU (chains are independent: weak scaling, no data-path collective).  Chains are
first burnt in (untimed) so the timed sweeps run on posterior-sized forests, not on root-only trees.

Launch: N=1 plain python; N>1 one rank per GPU under torch.distributed.run (NCCL only for the barrier / max).
Prints ONE JSON line on rank 0.  `--impl reference` times the reference algorithm's CPU port (oracle/) instead.
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

# dram__bytes_read.sum + dram__bytes_write.sum of sweep_trees_kernel, one `ncu --set full` capture of this workload
# (profiles/r1_e_ncu_full_summary.csv): the leaf-space state is L2-resident (86 % L2 hit rate), so DRAM traffic is
# far below the algorithmic bytes.
NCU_DRAM_BYTES_PER_LAUNCH = 2.268e9
NCU_SOURCE = "profiles/r1_e_ncu_full_summary.csv (ncu --set full, round 1)"

METRIC = "mcmc_proposals_per_sec_full_mll"
UNIT = "proposals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--m", type=int, default=200)
    ap.add_argument("--d", type=int, default=10)
    ap.add_argument("--chains-per-gpu", type=int, default=64)
    ap.add_argument("--burnin", type=int, default=120, help="untimed sweeps that bring the chains to posterior-sized forests")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--predict-candidates", type=int, default=1 << 21,
                    help="candidates of the predict figure in `extras` (default: BASELINE config 5's per-GPU shard, 2 Mi)")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0, help="target CPU seconds of the bounded baseline sample")
    return ap.parse_args()


def workload_name(a):
    return (f"BARK MCMC fit, synthetic TreeFunction data, N={a.n}, D={a.d} continuous, m={a.m} trees, "
            f"{a.chains_per_gpu} chains per GPU (BASELINE config 4)")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


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


# --------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference algorithm's CPU port (oracle/bark_oracle.py, numba + LAPACK)
# --------------------------------------------------------------------------------------------------------
def cpu_sample(a, forest, noise, scale, X, y, bounds, ft, budget_s, steps=1, warmup=0):
    """Time the dense K^-1 Woodbury sampler (the reference's algorithm) on ONE chain for bounded slices of a
    sweep.  Returns (proposals/s, description, list of per-step seconds)."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import bark_oracle as O
    p = O.BARKTrainParams(num_chains=1)
    # JIT warm-up on a tiny problem (excluded from timing)
    Xs, ys, bs, fs, _ = O.synthetic_problem(24, dim=a.d, m_true=4, seed=0)
    tiny = O.CpuChain(O.create_empty_forest(3), 0.1, 1.0, (Xs, ys), bs, fs, p)
    tiny.advance(0, 3, True)
    chain = O.CpuChain(forest, noise, scale, (X, y), bounds, ft, p)  # builds K^-1 (untimed, like chain init)
    m = forest.shape[0]
    # probe: 4 tree proposals to size the slice
    t0 = time.perf_counter(); chain.advance(0, min(4, m), False); probe = (time.perf_counter() - t0) / min(4, m)
    per_step = max(1, min(m, int(budget_s / max(probe, 1e-9) / max(steps + warmup, 1))))
    times, props, cur = [], 0, min(4, m) % m
    for it in range(warmup + steps):
        t_end = min(m, cur + per_step)
        hyper = t_end == m
        t0 = time.perf_counter()
        k = chain.advance(cur, t_end, hyper)
        dt = time.perf_counter() - t0
        cur = 0 if t_end == m else t_end
        if it >= warmup:
            times.append(dt); props += k
    desc = (f"1 chain, {props} proposals ({per_step} tree proposals per step, dense N x N Woodbury + full "
            f"refactorisation at the noise step) of the {a.n}x{a.m} workload on the host CPU; numba-jitted port of the "
            f"reference algorithm, BLAS threads = all cores; per-chain rate (chains are serial in the reference)")
    return props / sum(times), desc, times


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    from oracle import bark_oracle as O
    X, y, bounds, ft, _ = O.synthetic_problem(a.n, dim=a.d, m_true=50, seed=0)
    # start state: a few CPU sweeps are too slow at this size; use a prior-like forest of posterior size instead
    forest = O.create_empty_forest(a.m)
    rng = np.random.default_rng(0)
    cdf = np.cumsum([0.5, 0.1, 0.4])
    for _ in range(3):
        for t in range(a.m):
            new, lqp, st = O.get_tree_proposal(forest[t], bounds, ft, 0.95, 2.0, cdf, rng.random(5), 0)
            if np.isfinite(lqp) and (forest[t]["active"] & forest[t]["is_leaf"]).sum() < 4:
                forest[t] = new
    budget = max(20.0, min(200.0, 8.0 * (a.steps + a.warmup)))
    val, desc, times = cpu_sample(a, forest, 0.1, 1.0, X, y, bounds, ft, budget, steps=a.steps, warmup=a.warmup)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "step": "bounded slice of one chain's sweep on the host CPU"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------
def run_b200_arm(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import bark_b200 as B
    from bark_b200 import synthetic
    from bark_b200.sampler import ChainState

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    C, m, n = a.chains_per_gpu, a.m, a.n
    X, y, bounds, ft, _ = synthetic.synthetic_problem(n, dim=a.d, m_true=50, seed=0)
    params = B.BARKTrainParams(num_chains=C)
    forest0 = np.tile(B.create_empty_forest(m), (C, 1, 1))
    seed = 20261018
    st = ChainState(forest0, np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft, device=dev)
    chain_off = rank * C
    st.sweeps(params, a.burnin, seed, chain_offset=chain_off, sweep_offset=0)
    torch.cuda.synchronize()
    sweep_no = a.burnin

    # L2 flush between timed steps (rule: flush or exceed L2): write a 256 MiB buffer
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # ---- warm-up steps (untimed)
    for _ in range(a.warmup):
        st.sweeps(params, 1, seed, chain_offset=chain_off, sweep_offset=sweep_no)
        sweep_no += 1
    c_before = st.read()["counters"].cpu().numpy().astype(np.float64)

    # ---- timed: exactly K steps, each bracketed by CUDA events on the launching stream, L2 flushed before each
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for k in range(a.steps):
        flush_buf.zero_()
        ev[k][0].record()
        st.sweeps(params, 1, seed, chain_offset=chain_off, sweep_offset=sweep_no)
        ev[k][1].record()
        sweep_no += 1
    barrier()
    clk = clocks.stop()
    ms_steps = [e0.elapsed_time(e1) for e0, e1 in ev]
    ms_total = torch.tensor([float(sum(ms_steps))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    r = st.read()
    c_after = r["counters"].cpu().numpy().astype(np.float64)
    B.sampler.raise_for_status(r["status"].cpu().numpy())
    dc = (c_after - c_before).sum(axis=0)
    proposals_per_step = world * C * (m + 1)
    value = proposals_per_step * a.steps / (ms_total / 1e3)

    # ---- roofline of the dominant kernel (tree sweep): per-kernel CUDA-event times from the C ABI
    kt = max(4, min(a.steps, 10))
    cb = st.read()["counters"].cpu().numpy().astype(np.float64)
    ms_trees, ms_hyper = st.sweeps_timed(params, kt, seed, chain_offset=chain_off, sweep_offset=sweep_no)
    sweep_no += kt
    ca = st.read()["counters"].cpu().numpy().astype(np.float64)
    d2 = (ca - cb).sum(axis=0)
    # algorithmic bytes of the leaf-space formulation; only the lower triangle of B^-1 is kept:
    #   matvec evaluation: 4 B x extent^2   (read the lower triangle once)         -> counters[11] * 4
    #   accepted update  : 8 B x extent^2   (read + write the lower triangle once) -> counters[12] * 8
    #   v = Z^T u        : 4 B x wd x extent (leaf bitsets)                         -> counters[13] * wd * 4
    wd = ((n + 31) // 32 + 3) // 4 * 4
    alg_bytes = d2[11] * 4 + d2[12] * 8 + d2[13] * wd * 4
    peak, peak_src = measured_peaks()
    ach = alg_bytes / (ms_trees / 1e3) / 1e9
    acc_rate = d2[2] / max(d2[0], 1)
    # the reference's dense-state byte model: 8 N^2 (1 + a) bytes per tree proposal
    ref_model_gbs = (d2[0] * 8.0 * n * n * (1 + acc_rate)) / (ms_trees / 1e3) / 1e9
    p_used = r["p_used"].cpu().numpy()
    roofline = {
        "bound": "hbm", "kernel": "sweep_trees_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
        "frac": ach / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": NCU_SOURCE, "peak_source": peak_src,
        "ms_per_launch": ms_trees / kt, "share_of_step": ms_trees / (ms_trees + ms_hyper),
        "algorithmic_bytes_per_launch": alg_bytes / kt,
        "reference_dense_model": {"bytes_per_proposal": 8.0 * n * n * (1 + acc_rate), "equivalent_gbs": ref_model_gbs,
                                  "frac_of_hbm_peak": ref_model_gbs / peak,
                                  "note": "SURVEY 8d model of the reference's N x N Woodbury state; >1 because the "
                                          "leaf-space state is P x P (P ~ 2.3 m << N), lower triangle only, L2-resident"},
        "note": "achieved = leaf-space algorithmic bytes / CUDA-event time; the kernel is latency-bound (ncu: L2 hit 86 %, "
                "DRAM 3 % of peak), see DESIGN.md section 6",
        "hyper_kernel_ms_per_launch": ms_hyper / kt,
    }

    # ---- end to end through the public API (host buffers in, host samples out)
    e2e = None
    if True:
        host_forest = st.dforest.to_numpy()
        rr = st.read()
        h_noise, h_scale = rr["noise"].cpu().numpy(), rr["scale"].cpu().numpy()
        ke = max(2, min(a.steps, 50))  # one fit call of K sweeps (the API granularity: upload, state build, K sweeps, download)
        pe = B.BARKTrainParams(warmup_steps=0, num_samples=1, steps_per_sample=ke, num_chains=C)
        # one untimed call first: warms torch's pinned-host and device caching allocators (as the warm-up steps do
        # for the device-timed number)
        pw = B.BARKTrainParams(warmup_steps=0, num_samples=1, steps_per_sample=1, num_chains=C)
        B.run_bark_sampler((host_forest, h_noise, h_scale), (X, y), (bounds, ft), pw, seed=seed + 2,
                           chain_offset=chain_off, device=dev)
        barrier()
        t0 = time.perf_counter()
        ns, no, sc = B.run_bark_sampler((host_forest, h_noise, h_scale), (X, y), (bounds, ft), pe, seed=seed + 1,
                                        chain_offset=chain_off, device=dev)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = host_forest.nbytes + X.nbytes + y.nbytes + bounds.nbytes + ft.nbytes + h_noise.nbytes + h_scale.nbytes
        d2h = ns.nbytes + no.nbytes + sc.nbytes
        e2e = {"value": proposals_per_step * ke / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d / ke,
               "d2h_bytes_per_step": d2h / ke, "steps": ke,
               "what": "run_bark_sampler(host numpy forest/X/y -> host numpy samples), warm start, includes H2D, "
                       "state build (traversal, A, B^-1), sweeps, packing and D2H"}

    # ---- secondary figures (not the headline): posterior-predictive points/s and batched full-MLL evaluations/s
    extras = {}
    try:
        from bark_b200.mll import mll_batched_device
        from bark_b200.forest import gram_umma_device, traverse_device, forest_slots, _as_device_f64, _feat_types_device
        hf_all = st.dforest.to_numpy()
        rr = st.read()
        ps = B.PosteriorState((hf_all, rr["noise"].cpu().numpy(), rr["scale"].cpu().numpy()), (X, y), ft, a.d, device=dev)
        n_c = a.predict_candidates  # BASELINE config 5: 16 Mi candidates over 8 GPUs = 2 Mi per GPU, x 64 samples
        cand = torch.rand((n_c, a.d), dtype=torch.float64, device=dev)
        ps.predict_device(cand, mode=1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); ps.predict_device(cand, mode=1); e1.record(); torch.cuda.synchronize()
        ms_p = e0.elapsed_time(e1)
        extras["predict"] = {"points_per_s": n_c / ms_p * 1e3, "candidate_samples_per_s": n_c * C / ms_p * 1e3,
                             "n_candidates": n_c, "posterior_samples": C, "what": "mixture mean/var over all samples, "
                             "candidates resident in HBM, per GPU", "algorithmic_hbm_bytes_per_point": 8 * a.d + 16}
        # full log-MLL from scratch: traverse -> int8 tcgen05 Gram (fused FP64 epilogue) -> batched block LDL^T
        Xd, yd, ftd = _as_device_f64(X, dev), _as_device_f64(y.reshape(-1), dev), _feat_types_device(ft, dev)
        nz, sc_ = rr["noise"].to(torch.float64), rr["scale"].to(torch.float64)
        slots = forest_slots(hf_all)

        def full_mll():
            leaves = traverse_device(st.dforest, Xd, ftd)
            _, Kmat = gram_umma_device(leaves, leaves, slots=slots, want_counts=False, scale=sc_, noise=nz)
            return mll_batched_device(Kmat, yd)[0]
        full_mll(); torch.cuda.synchronize()
        ev3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev3[0].record(); leaves = traverse_device(st.dforest, Xd, ftd); ev3[1].record()
        _, Kmat = gram_umma_device(leaves, leaves, slots=slots, want_counts=False, scale=sc_, noise=nz); ev3[2].record()
        vals = mll_batched_device(Kmat, yd)[0]; ev3[3].record(); torch.cuda.synchronize()
        t_tr, t_gr, t_ml = ev3[0].elapsed_time(ev3[1]), ev3[1].elapsed_time(ev3[2]), ev3[2].elapsed_time(ev3[3])
        extras["full_mll_from_scratch"] = {
            "evals_per_s": C / (t_tr + t_gr + t_ml) * 1e3, "batch": C, "n": n, "m": m,
            "ms": {"traverse": t_tr, "gram_tcgen05_int8_incl_onehot_build": t_gr, "factorise_mll_fp64": t_ml},
            "gram_output_gbs": C * n * n * 8 / (t_gr / 1e3) / 1e9,
            "running_vs_scratch_max_rel_diff": float((vals - rr["mll"]).abs().div(rr["mll"].abs()).max().item())}
    except Exception as exc:
        extras["error"] = f"{type(exc).__name__}: {exc}"

    # ---- CPU baseline (rank 0, N=1 only): the reference algorithm's port on one chain, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            hf = st.dforest.to_numpy()[0]
            rr = st.read()
            val, desc, _ = cpu_sample(a, hf, float(rr["noise"][0]), float(rr["scale"][0]), X, y, bounds, ft, a.cpu_budget_s)
            cpu_baseline = {"value": val, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": desc}
        except Exception as exc:  # the baseline must never take the GPU number down with it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                            "sample": f"failed: {type(exc).__name__}: {exc}"}

    if rank == 0:
        ws_mb = C * (float(np.mean(((p_used + 15) // 16 * 16) ** 2)) * 8 + float(np.mean(p_used)) * wd * 4) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "chains_total": world * C, "burnin_sweeps": a.burnin,
                       "proposals_per_step": proposals_per_step,
                       "l2": f"L2 flushed (256 MiB write) before every timed step; hot state {ws_mb:.0f} MB per GPU",
                       "leaf_columns_per_chain_mean": float(np.mean(p_used)), "p_cap": st.p_cap,
                       "parallelism": f"chains sharded x{world}, no data-path collective"},
            "acceptance": {"tree_accept_rate": float(dc[2] / max(dc[0], 1)), "tree_valid_rate": float(dc[1] / max(dc[0], 1)),
                           "hyper_accept_rate": float(dc[4] / max(dc[3], 1))},
            "clocks": clk, "e2e": e2e, "gpu_launches": 3 * a.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
        return
    run_b200_arm(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
