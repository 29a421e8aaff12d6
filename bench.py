#!/usr/bin/env python
"""bench.py -- SAMTRON iteration throughput on the BASELINE.json stress configuration (C5).

One "step" = one full SAMTRON iteration (sample -> log-density / background -> Stein NG -> KL-constrained
component update -> trust-region weight update) of a K=512-component, D=256 full-covariance GMM on 65,536
samples per iteration (128 per component), GMM target, no sample reuse.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1|C2|C3|C3w|C4d|C4f|C5]
                    [--no-graph] [--no-parity] [--no-cpu] [--no-dense]

Prints ONE JSON line (rank 0).  The iteration runs as one CUDA graph per rank unless --no-graph (same kernels and NCCL
collectives, one launch).  At N > 1 the line carries `parity_vs_1gpu`: the sharded iteration against the single-GPU iteration
on identical samples, measured after the timed regions.  `--config` selects one of the other BASELINE.json configurations
(single GPU, through GmmviRunner, with an un-extrapolated CPU figure of the oracle beside it).  `value` = iterations/s with all inputs resident in HBM; `e2e` = the same through
GMMVI.train_iter with the step's noise coming from pinned host memory and the updated mixture read back;
`roofline` describes the dominant kernel (component log-density); `cpu_baseline` times the oracle (restated
reference, NumPy/OpenBLAS fp32) on a bounded sample of the same workload on this box's host cores.
`--impl reference` reports that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_COMP, DIM, PER_COMP = 512, 256, 128
TARGET_COMPONENTS = 10
# `roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, parsed from the newest
# committed `ncu --set full` summary of that kernel (profiles/r*_ncu_h16t_logdens.txt, written by profiles/ncu_summary.py).
# The captures are of the C5 shape (K=512, D=256, 65536 samples per launch); other shapes report null.
NCU_TRAFFIC_SHAPE = ("h16", 512, 256, 65536)
_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_dram_traffic():
    """-> (bytes per launch | None, source file | None) from the newest profiles/r*_ncu_h16t_logdens.txt."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_h16t_logdens.txt")))
    for path in reversed(files):
        vals = {}
        for line in open(path):
            parts = line.split()
            if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[1] in _UNIT:
                vals[parts[0]] = float(parts[2].replace(",", "")) * _UNIT[parts[1]]
        if len(vals) == 2:
            return sum(vals.values()), os.path.relpath(path, ROOT)
    return None, None
PRIOR_SCALE = 31.63          # configs/experiment_configs/gmm100.yml:11 (GMM-target experiments)


# ------------------------------------------------------------------------------------------------
def workload_arrays(K, D, seed=0, mean_scale=PRIOR_SCALE):
    """w = 1/K, mu_k ~ N(0, s^2 I), L_k = chol(A A^T / D + I); GMM target with 10
    components, means 100 (u - 0.5), covariance A^T A + I with A ~ 0.1 N(0, D) (target_distributions/gmm.py:135-145)."""
    rng = np.random.default_rng(seed)
    means = (rng.standard_normal((K, D)) * mean_scale).astype(np.float32)
    chols = np.empty((K, D, D), np.float32)
    for k in range(K):
        A = rng.standard_normal((D, D))
        chols[k] = np.linalg.cholesky(A @ A.T / D + np.eye(D)).astype(np.float32)
    tmeans = (100.0 * (rng.random((TARGET_COMPONENTS, D)) - 0.5)).astype(np.float32)
    tchols = np.empty((TARGET_COMPONENTS, D, D), np.float32)
    for j in range(TARGET_COMPONENTS):
        A = 0.1 * rng.normal(0.0, D, (D, D))
        tchols[j] = np.linalg.cholesky(A.T @ A + np.eye(D)).astype(np.float32)
    return means, chols, tmeans, tchols


def samtron_config(per_comp):
    return {
        "temperature": 1.0, "use_sample_database": False, "max_database_size": 10000000,
        "model_initialization": {"use_diagonal_covs": False, "prior_mean": 0.0, "initial_cov": 1.0},
        "ng_estimator_type": "Stein",
        "ng_estimator_config": {"only_use_own_samples": False, "use_self_normalized_importance_weights": True},
        "num_component_adapter_type": "fixed", "num_component_adapter_config": {},
        "sample_selector_type": "component-based",
        "sample_selector_config": {"desired_samples_per_component": per_comp, "ratio_reused_samples_to_desired": 0.0},
        "ng_based_updater_type": "trust-region", "ng_based_updater_config": {},
        "component_stepsize_adapter_type": "improvement-based",
        "component_stepsize_adapter_config": {"initial_stepsize": 0.1, "min_stepsize": 0.001, "max_stepsize": 1.0,
                                              "stepsize_inc_factor": 1.15, "stepsize_dec_factor": 0.85},
        "weight_updater_type": "trust-region", "weight_updater_config": {"use_self_normalized_importance_weights": True},
        "weight_stepsize_adapter_type": "improvement_based",
        "weight_stepsize_adapter_config": {"initial_stepsize": 1.0, "min_stepsize": 0.0001, "max_stepsize": 1.0,
                                           "stepsize_inc_factor": 1.15, "stepsize_dec_factor": 0.85},
    }


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (restated reference) on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_baseline(K, D, per_comp, sample_per_comp=4, seed=0):
    import oracle as O
    means, chols, tmeans, tchols = workload_arrays(K, D, seed)
    dt = np.float32
    g = O.OracleGMM(np.log(np.ones(K, dt) / K), means, chols, False, initial_stepsize=0.1)
    tgt = O.OracleGMM(np.log(np.ones(TARGET_COMPONENTS, dt) / TARGET_COMPONENTS), tmeans, tchols, False)

    def target(X):
        lq, gr, _ = O.log_density_and_grad(tgt, X)
        return lq, gr
    db = O.OracleSampleDB(D, False, False, None, dt)
    rng = np.random.default_rng(1)
    noise = lambda k, D_, n: rng.standard_normal((D_, n)).astype(dt)
    t0 = time.perf_counter()
    db._inv(chols)
    t_inv = time.perf_counter() - t0
    t0 = time.perf_counter()
    X, mapping, bg, lnpdfs, grads = O.vips_select_samples(g, db, target, sample_per_comp, 0.0, noise)
    t_sel = time.perf_counter() - t0
    t0 = time.perf_counter()
    H, gn = O.stein_ng(g, X, mapping, bg, lnpdfs, grads)
    t_stein = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.kl_constrained_update(g, H, gn, g.stepsizes, 1.0)
    t_upd = time.perf_counter() - t0
    t0 = time.perf_counter()
    elr = O.expected_log_ratios(g, X, bg, lnpdfs, 1.0, True)
    t_elr = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.trust_region_weight_update(g, elr, 1.0, 1.0)
    t_w = time.perf_counter() - t0
    scale = per_comp / sample_per_comp
    t_iter = (max(t_sel - t_inv, 0.0) + t_stein + t_elr) * scale + t_inv + t_upd + t_w
    n_s = K * sample_per_comp
    # log-density pairs/s of the reference-structured CPU path: one component_log_densities pass
    t0 = time.perf_counter()
    O.component_log_densities(g, X)
    t_ld = time.perf_counter() - t0
    return {
        "iters_per_sec": 1.0 / t_iter, "sec_per_iter": t_iter, "pairs_per_sec": n_s * K / t_ld,
        "stages_s": {"select+background": t_sel - t_inv, "stein": t_stein, "expected_log_ratios": t_elr,
                     "chol_inverse": t_inv, "kl_update": t_upd, "weight_update": t_w, "n_scale": scale},
        "sample": (f"oracle (restated reference, NumPy/OpenBLAS fp32) timed on N={n_s} of {K * per_comp} samples "
                   f"({sample_per_comp} of {per_comp} per component), full K={K} and D={D}; sample-proportional stages "
                   f"scaled x{scale:g}, Cholesky inverses and the KL/weight updates timed at full K"),
    }


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu_index = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------------
def shutdown_distributed(gmmvi_objects=()):
    """Release captured graphs (they hold NCCL kernels) before the process group goes away; a watchdog ends the process
    if the communicator teardown does not return (seen with graphs alive: the JSON line is already printed)."""
    import gc
    import torch
    import torch.distributed as dist
    for g in gmmvi_objects:
        try:
            g.enable_cuda_graph(False)
            g._graph_retired = None
        except Exception:
            pass
    gc.collect()
    torch.cuda.synchronize()
    if dist.is_available() and dist.is_initialized():
        sys.stdout.flush()
        t = threading.Timer(20.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gmmvi_b200 import ops, rng
    from gmmvi_b200.distributed import ShardContext
    from gmmvi_b200.experiments.target_distributions.gmm import GMM_LNPDF
    from gmmvi_b200.models.full_cov_gmm import FullCovGMM
    from gmmvi_b200.models.gmm_wrapper import GmmWrapper
    from gmmvi_b200.optimization.gmmvi import GMMVI

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    K, D, per = args.components, args.dim, args.per_comp
    N_total = K * per                            # samples per iteration (strong scaling: fixed, sharded over ranks)
    lo_hi = ShardContext(rank, world).row_range(N_total)
    N = lo_hi[1] - lo_hi[0]                      # rows this rank draws and evaluates

    def build(mean_scale, shard=True):
        means, chols, tmeans, tchols = workload_arrays(K, D, 0, mean_scale)
        model = FullCovGMM.from_cholesky(np.ones(K, np.float32) / K, means, chols, device=dev)
        tgt = GMM_LNPDF.from_cholesky(np.ones(TARGET_COMPONENTS) / TARGET_COMPONENTS, tmeans, tchols, device=dev)
        cfg = samtron_config(per)
        wrapper = GmmWrapper.build_from_config(model, cfg)
        g = GMMVI.build_from_config(cfg, tgt, wrapper)
        if world > 1 and shard:
            g.enable_sharding(ShardContext(rank, world))
        return g

    rng.set_seed(1234)
    gmmvi = build(PRIOR_SCALE)
    use_graph = not args.no_graph
    if use_graph:
        # one CUDA graph per iteration (the counterpart of the reference's tf.function around train_iter): the same
        # kernels and collectives, launched with one call instead of ~170 -- at 8 GPUs a rank's kernels take ~4 ms and
        # the Python between the launches would otherwise set the pace
        gmmvi.enable_cuda_graph()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- HBM-resident timing --------------------------------------------------------------------
    for _ in range(args.warmup):
        gmmvi.train_iter()
    sync_all()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = ops.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        gmmvi.train_iter()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = ops.kernel_launches() - l0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    finite = bool(torch.isfinite(gmmvi.model.means).all() and torch.isfinite(gmmvi.model.chol_cov).all())
    kl_evals = ops.last_update_evals.float() if ops.last_update_evals is not None else None
    kl_evals_stats = None if kl_evals is None else {"mean": float(kl_evals.mean()), "max": float(kl_evals.max()),
                                                    "success": float(gmmvi.ng_based_updater.last_success.float().mean())}

    # ---- end to end: host noise in, updated mixture out -------------------------------------------
    # Every step copies its own noise shard from pinned host memory and (rank 0) reads the updated mixture back;
    # the copies run on two copy streams so that step i+1's upload and step i's read-back overlap the kernels of
    # their neighbours (double-buffered noise; the result is read back from a per-step snapshot).  All of it is inside
    # the timed region.
    hostE = [torch.empty((N, D), dtype=torch.float32).pin_memory().normal_() for _ in range(2)]
    devE = [torch.empty((N, D), dtype=torch.float32, device=dev) for _ in range(2)]
    read_back = rank == 0
    out_w = torch.empty(K, dtype=torch.float32).pin_memory()
    out_m = torch.empty((K, D), dtype=torch.float32).pin_memory()
    out_c = torch.empty((K, D, D), dtype=torch.float32).pin_memory()
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def e2e_run(steps):
        cur = torch.cuda.current_stream(dev)
        up = [torch.cuda.Event() for _ in range(steps)]
        done = [torch.cuda.Event() for _ in range(steps)]

        def upload(i):
            with torch.cuda.stream(s_h2d):
                if i >= 2:
                    s_h2d.wait_event(done[i - 2])          # the buffer's previous reader
                devE[i % 2].copy_(hostE[i % 2], non_blocking=True)
                up[i].record(s_h2d)
        upload(0)
        for i in range(steps):
            if i + 1 < steps:
                upload(i + 1)
            cur.wait_event(up[i])
            gmmvi.train_iter(noise=devE[i % 2])
            if read_back:
                # a graph replay updates the mixture in place (static buffers): read back from a snapshot taken on the
                # compute stream (134 MB device copy, ~0.05 ms) so that the next step does not race the D2H copy
                lw, mu, ch = (t.clone() for t in (gmmvi.model.log_weights, gmmvi.model.means, gmmvi.model.chol_cov))
            done[i].record(cur)
            if read_back:
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(done[i])
                    out_w.copy_(lw, non_blocking=True)
                    out_m.copy_(mu, non_blocking=True)
                    out_c.copy_(ch, non_blocking=True)
                    for t_ in (lw, mu, ch):
                        t_.record_stream(s_d2h)
        torch.cuda.synchronize()
    e2e_run(2)
    sync_all()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    # ---- N > 1: the sharded iteration against the single-GPU iteration on the same samples -------------------
    # (outside every timed region)  The counter-based generator draws a sample from its GLOBAL row index, so a fresh
    # mixture iterated `parity_iters` times from a fixed seed sees the same samples sharded over `world` ranks as on one
    # GPU; rank 0 then repeats the run unsharded and compares the resulting mixtures (max-abs / max-abs per tensor).
    parity = None
    if world > 1 and not args.no_parity:
        rng.set_seed(4321)
        gs = build(PRIOR_SCALE)
        for _ in range(args.parity_iters):
            gs.train_iter()
        sh = [t.detach().clone() for t in (gs.model.means, gs.model.chol_cov, gs.model.log_weights)]
        same = torch.tensor([1.0], device=dev)
        for t_ in sh:                       # every rank must hold the same replicated mixture, bit for bit
            ref_t = t_.clone()
            dist.broadcast(ref_t, src=0)
            if not torch.equal(ref_t, t_):
                same.zero_()
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        del gs
        sync_all()
        if rank == 0:
            rng.set_seed(4321)
            g1 = build(PRIOR_SCALE, shard=False)
            for _ in range(args.parity_iters):
                g1.train_iter()
            torch.cuda.synchronize()
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
            parity = {"iterations": args.parity_iters, "means": rel(sh[0], g1.model.means),
                      "chol": rel(sh[1], g1.model.chol_cov),
                      "weights": rel(torch.exp(sh[2]), g1.model.weights),
                      "ranks_bit_identical": bool(same.item() == 1.0),
                      "tolerance": 1e-4,
                      "note": "sharded vs single-GPU mixture after the same iterations from the same seed (identical "
                              "samples: counter-based generator keyed on the global row); sums are re-associated "
                              "across ranks, nothing else differs"}
            del g1
        sync_all()

    # ---- dominant kernel: component log-density, timed alone with CUDA events -------------------------
    linv, prec, cst = gmmvi.model.prepared()
    X = gmmvi.sample_db.samples.contiguous()
    reps = 5
    ops.logdens_full(X, gmmvi.model.means, linv, cst, memo=False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.logdens_full(X, gmmvi.model.means, linv, cst, memo=False)
    e1.record()
    torch.cuda.synchronize()
    ld_ms = e0.elapsed_time(e1) / reps
    pairs = float(X.shape[0]) * K
    flop_alg = pairs * (D * D + 4 * D)
    hbm, bf16, bf16_sus, how = measured_peaks()
    kind = ops.logdens_kernel_kind(D)
    # the fp16-split kernel runs kind::f16 MMAs (peak = the measured bf16/fp16 rate); the TF32 kernel half of that
    tc_peak = bf16 if kind == "h16" else bf16 / 2.0
    peak_note = (f"{how} bf16 burst {bf16} TFLOP/s (kind::f16 MMAs run at the bf16 rate)" if kind == "h16" else
                 f"{how} bf16 burst {bf16} TFLOP/s / 2 (TF32 dense rate is half the bf16 rate)")
    achieved = flop_alg / (ld_ms * 1e-3) / 1e12
    traffic, traffic_src = ncu_dram_traffic() if (kind, K, D, int(X.shape[0])) == NCU_TRAFFIC_SHAPE else (None, None)
    # executed tensor work of the split-precision kernel: 3 MMAs per 16-column step with N = Dp - 16 jb (h16)
    if kind == "h16":
        Dp = (D + 63) // 64 * 64
        mma_flop_pair = 3 * 2 * 16 * sum(Dp - 16 * jb for jb in range(Dp // 16))
    else:
        mma_flop_pair = 3 * 2 * 32 * sum(D - 32 * kb for kb in range(D // 32)) if D % 32 == 0 else None

    # ---- dense variant: overlapping components (nothing can be skipped) -------------------------------
    dense = None
    if rank == 0 and world == 1 and not args.no_dense:
        g2 = build(0.05)
        for _ in range(2):
            g2.train_iter()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(max(2, args.steps // 2)):
            g2.train_iter()
        e1.record()
        torch.cuda.synchronize()
        dense = max(2, args.steps // 2) / (e0.elapsed_time(e1) * 1e-3)
        del g2
    dense_block = None
    if dense is not None:
        # algorithmic flops of the full iteration with nothing skipped,
        # P [(D^2 + 4D) 3 + (D^2 + 3D) + 2 D^2] + N (D^2 + D) + K n_kl 3 D^3 + K D^3  (~ 6 P D^2 = 13.3 TFLOP at C5)
        P_pairs = float(N_total) * K
        n_kl = float(kl_evals.mean()) if kl_evals is not None else 10.0
        dense_flop = (P_pairs * ((D * D + 4 * D) * 3 + (D * D + 3 * D) + 2 * D * D) + N_total * (D * D + D)
                      + K * n_kl * 3.0 * D ** 3 + K * float(D) ** 3)
        hbm_, bf16_, _, how_ = measured_peaks()
        dense_block = {"iterations_per_sec": dense, "ms_per_step": 1e3 / dense,
                       "algorithmic_tflop_per_iteration": dense_flop / 1e12,
                       "achieved_tflops": dense_flop * dense / 1e12, "peak_tflops": bf16_,
                       "frac": dense_flop * dense / 1e12 / bf16_,
                       "note": "components overlap (mean scale 0.05): no (component, sample block) is skipped anywhere; "
                               "whole-iteration algorithmic flops over the whole-iteration time against the " + how_ +
                               " bf16 burst; per-kernel ncu figures: profiles/r02_ncu_stein_tc_full_flush32.txt, "
                               "r02_ncu_mixgrad_h16_dense.txt, r02_ncu_gsum2.txt"}

    if rank != 0:
        shutdown_distributed([gmmvi])
        return
    cpu = cpu_baseline(K, D, per, args.cpu_sample_per_comp) if (world == 1 and not args.no_cpu) else None
    line = {
        "metric": "samtron_iterations_per_sec", "value": args.steps / (ms_total * 1e-3),
        "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C5 stress: SAMTRON (Stein NG + KL-constrained update + trust-region weights), "
                               f"K={K} full-cov components, D={D}, {per} samples/component "
                               f"({N_total} samples/iteration sharded over {world} GPU), GMM target "
                               f"({TARGET_COMPONENTS} comps)",
                   "samples_per_iteration": N_total, "components": K, "dim": D,
                   "parallelism": f"samples sharded over {world} rank(s), NCCL all-reduce of per-component statistics, "
                                  f"component update sharded + all-gather",
                   "l2": "working set per step (~1.2 GB: [K,N] densities, [K,D,D] factors) exceeds the 126 MB L2",
                   "pairs_per_sec_full_iteration": N_total * K / (ms_per_step * 1e-3),
                   "dense_variant_iterations_per_sec": dense, "dense_variant": dense_block, "finite": finite,
                   "cuda_graph": use_graph,
                   "kl_evaluations_per_component": kl_evals_stats},
        "logdens_pairs_per_sec": pairs / (ld_ms * 1e-3),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": achieved / tc_peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes": 4.0 * (X.shape[0] * D + 2 * K * D * D // 2 + K * X.shape[0]),
                     "kernel": ops.logdens_kernel_name(D), "launch_ms": ld_ms,
                     "algorithmic_flop_per_pair": D * D + 4 * D,
                     "executed_mma_flop_per_pair": mma_flop_pair,
                     "executed_mma_frac": (None if mma_flop_pair is None else
                                           pairs * mma_flop_pair / (ld_ms * 1e-3) / 1e12 / tc_peak),
                     "peak_source": peak_note},
        "e2e": {"value": args.steps / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": N_total * D * 4,
                "d2h_bytes_per_step": (K + K * D + K * D * D) * 4,
                "note": "per step: every rank uploads its noise shard from pinned host memory, rank 0 reads the "
                        "updated mixture (log-weights, means, Cholesky factors) back; copies on separate streams, "
                        "overlapping the neighbouring steps' kernels, all inside the timed region"},
        "gpu_launches": launches, "clocks": clk,
    }
    if parity is not None:
        line["parity_vs_1gpu"] = parity
    if cpu is not None:
        line["cpu_baseline"] = {"value": cpu["iters_per_sec"], "unit": "iterations/s", "cores": os.cpu_count(),
                                "kind": "port", "sample": cpu["sample"], "pairs_per_sec": cpu["pairs_per_sec"],
                                "stages_s": cpu["stages_s"], "extrapolated": True,
                                "extrapolation": f"sample-proportional stages timed on 1/{cpu['stages_s']['n_scale']:g} "
                                                 "of the samples and scaled linearly"}
    print(json.dumps(line), flush=True)
    shutdown_distributed([gmmvi])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, D, per = args.components, args.dim, args.per_comp
    world = int(os.environ.get("WORLD_SIZE", "1"))
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        vals.append(cpu_baseline(K, D, per, args.cpu_sample_per_comp))
    best = max(vals, key=lambda c: c["iters_per_sec"])
    # strong scaling: the iteration's total work does not depend on the number of GPUs of the other arm
    sec = best["sec_per_iter"]
    value = 1.0 / sec
    line = {
        "impl": "reference", "metric": "samtron_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C5 stress on host CPU: K={K}, D={D}, {per * K} samples/iteration",
                   "samples_per_iteration": per * K, "components": K, "dim": D},
        "extrapolated": True,
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": best["sample"], "extrapolated": True,
                         "extrapolation": f"sample-proportional stages timed on 1/{best['stages_s']['n_scale']:g} of the "
                                          "samples and scaled linearly (BASELINE.md section 3): a stated estimate, not "
                                          "a measurement of the full configuration"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# The other BASELINE.json configurations (C1 - C4) through the reference-facing runner: python bench.py --config C1
# ------------------------------------------------------------------------------------------------
CONFIG_DOC = {
    "C1": "examples/5_samtron_20D_student-T.py: SAMTRON, 20-D Student-t mixture target, 45 initial components (+1 / 60 "
          "iterations), 200 samples per component, no reuse",
    "C2": "examples/6_samtron_planar4.py: SAMTRON, 10-link planar robot with 4 goals, 100 initial components (+1 per "
          "iteration, del_iters 10), 100 samples per component",
    "C3": "synthetic 100-D 10-mode GMM target, 50 full-covariance components, 4096 samples / iteration (mixture-based "
          "selector), MORE + trust-region weights; N < F = 5151 features: rank deficient, as SURVEY.md warns",
    "C3w": "C3 with 12288 samples / iteration (N > 2 F: a well-posed regression)",
    "C4d": "synthetic 200-D Student-t mixture, 256 DIAGONAL components, 64 samples per component (16384 / iteration), "
           "Stein + iBLR, fixed stepsize 1e-4",
    "C4f": "C4 with full covariances",
}


def build_config_runner(name, use_graph=True):
    from gmmvi_b200.configs import get_default_algorithm_config, get_default_experiment_config, update_config
    from gmmvi_b200.gmmvi_runner import GmmviRunner
    rc = {"log_metrics_interval": 10 ** 9, "use_cuda_graph": use_graph}
    if name in ("C1", "C2"):
        if name == "C1":
            exp, over = "stm20", {"sample_selector_config": {"desired_samples_per_component": 200,
                                                            "ratio_reused_samples_to_desired": 0.0},
                                  "model_initialization": {"num_initial_components": 45}}
        else:
            exp, over = "planar_robot_4", {"num_component_adapter_config": {"del_iters": 10, "add_iters": 1},
                                           "sample_selector_config": {"desired_samples_per_component": 100,
                                                                      "ratio_reused_samples_to_desired": 0.0},
                                           "model_initialization": {"num_initial_components": 100}}
        algo = update_config(get_default_algorithm_config("SAMTRON"), over)
        config = update_config(update_config(get_default_experiment_config(exp), {"start_seed": 1}), algo)
        config["gmmvi_runner_config"] = rc
        return GmmviRunner.build_from_config(config), config
    from gmmvi_b200 import rng as grng
    grng.set_seed(1)
    if name in ("C3", "C3w"):
        from gmmvi_b200.experiments.target_distributions.gmm import make_target
        D, K, target, codeword, diag, prior_scale, initial_cov = 100, 50, make_target(100), "ZEPTFOX", False, 31.63, 1.0
        over = {"sample_selector_config": {"desired_samples_per_component": 4096 if name == "C3" else 12288,
                                           "ratio_reused_samples_to_desired": 0.0},
                "component_stepsize_adapter_config": {"initial_stepsize": 0.01}}
    else:
        from gmmvi_b200.experiments.target_distributions.student_t_mixture import make_target
        D, K, target, codeword, diag, prior_scale, initial_cov = (200, 256, make_target(200, False, device="cuda"), "SEMYFUX",
                                                                  name == "C4d", 100.0, 300.0)
        over = {"sample_selector_config": {"desired_samples_per_component": 64, "ratio_reused_samples_to_desired": 0.0},
                "component_stepsize_adapter_config": {"initial_stepsize": 1e-4}}
    algo = update_config(get_default_algorithm_config(codeword), over)
    config = update_config({"start_seed": 1, "use_sample_database": False, "max_database_size": 10000000, "temperature": 1.0,
                            "model_initialization": {"use_diagonal_covs": diag, "num_initial_components": K,
                                                     "prior_mean": 0.0, "prior_scale": prior_scale,
                                                     "initial_cov": initial_cov},
                            "gmmvi_runner_config": rc}, algo)
    config["target_fn"] = target
    return GmmviRunner.build_from_config(config), config


def cpu_baseline_config(name, runner, config, iters):
    """The oracle (restated reference, NumPy fp32) iterating the SAME configuration from the device's current mixture on
    the host cores: 1 warm-up + `iters` timed iterations, median.  Nothing is extrapolated at these sizes."""
    import oracle as O
    dt = np.float32
    g = runner.gmmvi
    m = g.model
    diag = bool(m.diagonal_covs)
    f = lambda t: t.detach().cpu().numpy().astype(dt)
    nca = config["num_component_adapter_config"]
    H = 10000 if "del_iters" not in nca else max(2 * max(2, nca["del_iters"]) + 8, 64)     # runner: 10000 (quirk 17); a
    # shorter window that still covers del_iters keeps the host copy of the history from dominating the CPU figure
    og = O.OracleGMM(f(m.log_weights), f(m.means), f(m.chol_cov), diag,
                     initial_stepsize=float(config["component_stepsize_adapter_config"]["initial_stepsize"]),
                     initial_regularizer=float(config["ng_estimator_config"].get("initial_l2_regularizer", 1e-12)),
                     max_reward_history_length=H)
    og.stepsizes = f(m.stepsizes)
    tgt = g.sample_selector.target_distribution
    tname = type(tgt).__name__
    if tname == "StudentTMixture_LNPDF":
        target = O.student_t_mixture_target(tgt.target_weights.numpy(), tgt.target_means.numpy(), tgt.target_covs.numpy(),
                                            tgt.alpha, dt)
    elif tname == "PlanarRobot":
        target = O.planar_robot_target(tgt._num_dimensions, tgt._num_goals, dt=dt)
    else:
        target = O.gmm_target(tgt.target_weights.numpy(), tgt.target_means.numpy(), tgt.target_covs.numpy(), dt)
    D = m.num_dimensions
    keep = bool(config["use_sample_database"])
    db = O.OracleSampleDB(D, diag, keep, config["max_database_size"] if keep else None, dt)
    sel = config["sample_selector_config"]
    cs_type = config["component_stepsize_adapter_type"]
    cs_cfg = {k: v for k, v in config["component_stepsize_adapter_config"].items() if k != "initial_stepsize"}
    if cs_type == "decaying":
        cs_cfg["initial_stepsize"] = config["component_stepsize_adapter_config"]["initial_stepsize"]
    ws_cfg = config["weight_stepsize_adapter_config"]
    wad = None
    if config["weight_stepsize_adapter_type"] == "improvement_based":
        wad = O.ImprovementBasedWeightStepsize(ws_cfg["initial_stepsize"], ws_cfg["min_stepsize"], ws_cfg["max_stepsize"],
                                               ws_cfg["stepsize_inc_factor"], ws_cfg["stepsize_dec_factor"], dt=dt)
    elif config["weight_stepsize_adapter_type"] == "decaying":
        wad = O.DecayingWeightStepsize(ws_cfg["initial_stepsize"], ws_cfg["annealing_exponent"], dt=dt)
    cfg = O.IterationConfig(
        sample_selector=config["sample_selector_type"], desired_samples_per_component=sel["desired_samples_per_component"],
        ratio_reused_samples_to_desired=sel["ratio_reused_samples_to_desired"], ng_estimator=config["ng_estimator_type"],
        only_use_own_samples=config["ng_estimator_config"]["only_use_own_samples"],
        ng_self_normalized=config["ng_estimator_config"]["use_self_normalized_importance_weights"],
        updater=config["ng_based_updater_type"], component_stepsize=cs_type,
        component_stepsize_cfg={"min_stepsize": cs_cfg.get("min_stepsize"), "max_stepsize": cs_cfg.get("max_stepsize"),
                                "inc": cs_cfg.get("stepsize_inc_factor"), "dec": cs_cfg.get("stepsize_dec_factor")}
        if cs_type == "improvement-based" else cs_cfg,
        weight_updater=config["weight_updater_type"],
        weight_self_normalized=config["weight_updater_config"]["use_self_normalized_importance_weights"],
        weight_stepsize=float(ws_cfg["initial_stepsize"]), temperature=float(config["temperature"]))
    adapter = None
    if config["num_component_adapter_type"] == "adaptive":
        a = {k: nca[k] for k in ("del_iters", "add_iters", "max_components", "thresholds_for_add_heuristic",
                                 "min_weight_for_del_heuristic", "num_database_samples")}
        mi = config["model_initialization"]
        adapter = O.VipsComponentAdaptation(og, db, mi["prior_mean"], mi["initial_cov"], **a)
    rs = np.random.default_rng(0)
    noise_fn = lambda k, D_, n: rs.standard_normal((D_, n)).astype(dt)
    uniform_fn = lambda n: rs.uniform(size=n).astype(dt)
    times = []
    for it in range(iters + 1):
        t0 = time.perf_counter()
        O.train_iter(og, db, target, cfg, noise_fn, wad, uniform_fn)
        if adapter is not None:
            adapter.adapt_number_of_components(g.num_updates + it + 1, lambda: float(rs.uniform()),
                                               lambda n: rs.permutation(n), target)
        times.append(time.perf_counter() - t0)
    sec = float(np.median(times[1:]))
    return {"value": 1.0 / sec, "unit": "iterations/s", "cores": os.cpu_count(), "kind": "port", "extrapolated": False,
            "sample": f"oracle (restated reference, NumPy/OpenBLAS fp32), the full configuration from the device's current "
                      f"mixture (K={og.num_components}): 1 warm-up + median of {iters} iteration(s) = {sec * 1e3:.1f} ms",
            "sec_per_iter": sec}


def run_config(args):
    import torch
    from gmmvi_b200 import ops
    name = args.config
    torch.cuda.set_device(0)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):          # the config helpers print like the reference's; stdout = ONE JSON line
        runner, config = build_config_runner(name, not args.no_graph)
    g = runner.gmmvi
    warm = max(args.warmup, 12)              # past the first component additions / the deletion window of C1 / C2
    if name in ("C1", "C2") and args.steps < 120:
        args.steps = 120                     # sub-millisecond iterations: span two component additions of C1 (every 60)
    for n in range(warm):
        g.train_iter()
    torch.cuda.synchronize()
    captures0 = g.graph_captures
    clocks = ClockSampler(0)
    clocks.start()
    l0 = ops.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g.train_iter()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ops.kernel_launches() - l0
    captures_timed = g.graph_captures - captures0
    clk = clocks.stop()
    # end to end: the reference-facing call of the examples, GmmviRunner.iterate_and_log (a device synchronisation and the
    # read-back of the cheap metrics every iteration)
    t0 = time.perf_counter()
    for n in range(args.steps):
        out = runner.iterate_and_log(warm + args.steps + n)
    e2e_s = time.perf_counter() - t0
    m = g.model
    K, D = m.num_components, m.num_dimensions
    N = int(g.sample_db.samples.shape[0]) if not g.sample_db.keep_samples else None
    sel = config["sample_selector_config"]
    n_iter = sel["desired_samples_per_component"] * (K if config["sample_selector_type"] == "component-based" else 1)
    finite = bool(torch.isfinite(m.means).all() and torch.isfinite(m.chol_cov).all())
    # dominant sample x component kernel alone: one component_log_densities pass over an iteration's worth of samples
    X = m.sample(n_iter)[0].contiguous()
    m.component_log_densities(X)
    torch.cuda.synchronize()
    ops.clear_caches()
    reps = 10
    e0.record()
    for _ in range(reps):
        if m.diagonal_covs:
            ops.logdens_diag(X, m.means, m.chol_cov)
        else:
            linv, _, cst = m.prepared(need_prec=False)
            ops.logdens_full(X, m.means, linv, cst, memo=False)
    e1.record()
    torch.cuda.synchronize()
    ld_ms = e0.elapsed_time(e1) / reps
    pairs = float(n_iter) * K
    hbm, bf16, _, how = measured_peaks()
    if m.diagonal_covs or D <= 64:
        # bandwidth-shaped: SURVEY.md section 8(d) bytes model 4 (N D + 2 K D + K + K N) (full: K D^2 / 2 factor entries)
        par = 2 * K * D if m.diagonal_covs else K * D * (D + 1) // 2 + K * D
        alg_bytes = 4.0 * (n_iter * D + par + K + K * n_iter)
        ach = alg_bytes / (ld_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                "algorithmic_bytes": alg_bytes, "peak_source": f"{how} HBM copy bandwidth"}
    else:
        flop = pairs * (D * D + 4 * D)
        ach = flop / (ld_ms * 1e-3) / 1e12
        kind = ops.logdens_kernel_kind(D)
        peak = bf16 if kind == "h16" else 63.98       # SIMT fp32: measured cuBLAS fp32 (profiles/r02_measured_peaks_tf32_fp64.json)
        roof = {"bound": "tensor" if kind == "h16" else "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": None, "algorithmic_flop_per_pair": D * D + 4 * D,
                "peak_source": f"{how} bf16 burst" if kind == "h16" else "measured cuBLAS fp32 SIMT GEMM"}
    roof.update({"kernel": "component log-density: " + ("gvi::logdens_diag_kernel" if m.diagonal_covs else
                                                         ("gvi::sd::logdens_small_kernel" if D <= 32 else ops.logdens_kernel_name(D))),
                 "launch_ms": ld_ms, "pairs_per_launch": pairs})
    cpu = None
    if not args.no_cpu:
        try:
            cpu = cpu_baseline_config(name, runner, config, 3 if name in ("C1", "C2") else 1)
        except Exception as e:            # e.g. C3: N < F makes the reference's normal equations singular (LinAlgError)
            cpu = {"value": None, "unit": "iterations/s", "cores": os.cpu_count(), "kind": "port",
                   "error": f"{type(e).__name__}: {e}"}
    line = {"metric": "gmmvi_iterations_per_sec", "value": args.steps / (ms * 1e-3), "unit": "iterations/s", "n_gpus": 1,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{name}: {CONFIG_DOC[name]}", "components": K, "dim": D,
                       "samples_per_iteration": n_iter, "finite": finite,
                       "cuda_graph": bool(g._graph_enabled and g._graph), "graph_captures_in_timed_region": captures_timed,
                       "l2": "iteration working set is rewritten every step; no buffer is reused across timed steps"},
            "logdens_pairs_per_sec": pairs / (ld_ms * 1e-3), "roofline": roof,
            "e2e": {"value": args.steps / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 8,
                    "note": "GmmviRunner.iterate_and_log: the call of the reference's examples; the samples are drawn on the "
                            "device by the algorithm itself (no host input), every iteration ends with a device "
                            "synchronisation and the read-back of the cheap metrics"},
            "gpu_launches": launches, "clocks": clk}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--components", type=int, default=K_COMP)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--per-comp", type=int, default=PER_COMP)
    ap.add_argument("--cpu-sample-per-comp", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--config", default="C5", choices=["C1", "C2", "C3", "C3w", "C4d", "C4f", "C5"],
                    help="BASELINE.json configuration; C5 (default) is the headline stress configuration")
    ap.add_argument("--no-dense", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the iteration kernel by kernel instead of as a CUDA graph")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-vs-single-GPU check at N > 1")
    ap.add_argument("--parity-iters", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "C5":
        run_config(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
