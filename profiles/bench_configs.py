"""Iteration times of the BASELINE.json configurations C1-C4 through the reference-facing runner / module API
(C5 is bench.py).  usage: python profiles/bench_configs.py [iterations] [C1,C3,...] [--kernels]
--kernels adds the per-kernel device time of the timed iterations (torch.profiler, a second pass)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmmvi_b200.configs import get_default_algorithm_config, get_default_experiment_config, update_config  # noqa: E402
from gmmvi_b200.gmmvi_runner import GmmviRunner  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
iters = int(args[0]) if len(args) > 0 else 60
only = set(args[1].split(",")) if len(args) > 1 else None
kernels = "--kernels" in sys.argv


def timed(name, make_runner, warm=10):
    if only is not None and name.split()[0] not in only:
        return
    runner = make_runner()
    for n in range(warm):
        runner.iterate_and_log(n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for n in range(warm, warm + iters):
        runner.iterate_and_log(n)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    m = runner.gmmvi.model
    print(f"{name:58s} {dt * 1e3:8.2f} ms/iter  K={m.num_components:4d}  finite={bool(torch.isfinite(m.means).all())}", flush=True)
    if kernels:
        n_prof = min(iters, 5)
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for n in range(warm + iters, warm + iters + n_prof):
                runner.iterate_and_log(n)
            torch.cuda.synchronize()
        for e in sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)[:10]:
            print(f"    {e.self_device_time_total / n_prof / 1e3:9.3f} ms  n={e.count / n_prof:7.1f}  {e.key[:100]}", flush=True)


def runner_for(experiment, codeword, overrides):
    algo = update_config(get_default_algorithm_config(codeword), overrides)
    env = update_config(get_default_experiment_config(experiment), {"start_seed": 1})
    config = update_config(env, algo)
    config["gmmvi_runner_config"] = {"log_metrics_interval": 10 ** 9}
    return GmmviRunner.build_from_config(config)


# C1: examples/5_samtron_20D_student-T.py (D=20, 45 initial components, 200 samples per component)
timed("C1 SAMTRON stm20 (ex.5)", lambda: runner_for("stm20", "SAMTRON", {
    "sample_selector_config": {"desired_samples_per_component": 200, "ratio_reused_samples_to_desired": 0.0},
    "model_initialization": {"num_initial_components": 45}}))
# C2: examples/6_samtron_planar4.py (D=10, 100 initial components, 100 samples per component)
timed("C2 SAMTRON planar_robot_4 (ex.6)", lambda: runner_for("planar_robot_4", "SAMTRON", {
    "num_component_adapter_config": {"del_iters": 10, "add_iters": 1},
    "sample_selector_config": {"desired_samples_per_component": 100, "ratio_reused_samples_to_desired": 0.0},
    "model_initialization": {"num_initial_components": 100}}))


def direct_runner(name, D, K, target, codeword, overrides, diag=False, prior_scale=31.63, initial_cov=1.0):
    """C3 / C4: synthetic fixed-K configurations of BASELINE.json built through the same runner (target passed as
    config['target_fn'], setup_experiment.py:14-16)."""
    algo = update_config(get_default_algorithm_config(codeword), overrides)
    config = update_config({"start_seed": 1, "use_sample_database": False, "max_database_size": 10000000, "temperature": 1.0,
                            "model_initialization": {"use_diagonal_covs": diag, "num_initial_components": K,
                                                     "prior_mean": 0.0, "prior_scale": prior_scale,
                                                     "initial_cov": initial_cov},
                            "gmmvi_runner_config": {"log_metrics_interval": 10 ** 9}}, algo)
    config["target_fn"] = target
    timed(name, lambda: GmmviRunner.build_from_config(config))


from gmmvi_b200.experiments.target_distributions.gmm import make_target as make_gmm_target  # noqa: E402
from gmmvi_b200.experiments.target_distributions.student_t_mixture import make_target as make_stm_target  # noqa: E402

# C3: 100-D 10-mode GMM target, 50 full-covariance components, 4096 samples per iteration (mixture-based selector:
# desired_samples_per_component is the TOTAL, sample_selector.py:322), MORE + trust-region weights
direct_runner("C3 GMM100 target, K=50, N=4096, MORE + TR weights", 100, 50, make_gmm_target(100), "ZEPTFOX", {
    "sample_selector_config": {"desired_samples_per_component": 4096, "ratio_reused_samples_to_desired": 0.0},
    "component_stepsize_adapter_config": {"initial_stepsize": 0.01}})
# C4: 200-D Student-T mixture, 256 components, 64 samples per component (16384 per iteration), Stein + iBLR, fixed
# stepsize 1e-4 (examples/3: guidance for iBLR), diagonal and full covariances
for diag in (True, False):
    direct_runner(f"C4 STM200 target, K=256, N=16384, Stein + iBLR, {'diagonal' if diag else 'full'} cov", 200, 256,
                  make_stm_target(200, False, device="cuda"), "SEMYFUX", {
        "sample_selector_config": {"desired_samples_per_component": 64, "ratio_reused_samples_to_desired": 0.0},
        "component_stepsize_adapter_config": {"initial_stepsize": 1e-4}}, diag=diag, prior_scale=100.0, initial_cov=300.0)
