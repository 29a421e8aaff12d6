"""Iteration times of the BASELINE.json configurations C1-C4 through the reference-facing runner / module API
(C5 is bench.py).  usage: python profiles/bench_configs.py [iterations]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmmvi_b200.configs import get_default_algorithm_config, get_default_experiment_config, update_config  # noqa: E402
from gmmvi_b200.gmmvi_runner import GmmviRunner  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60


def timed(name, runner, warm=10):
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


def runner_for(experiment, codeword, overrides):
    algo = update_config(get_default_algorithm_config(codeword), overrides)
    env = update_config(get_default_experiment_config(experiment), {"start_seed": 1})
    config = update_config(env, algo)
    config["gmmvi_runner_config"] = {"log_metrics_interval": 10 ** 9}
    return GmmviRunner.build_from_config(config)


# C1: examples/5_samtron_20D_student-T.py (D=20, 45 initial components, 200 samples per component)
timed("C1 SAMTRON stm20 (ex.5)", runner_for("stm20", "SAMTRON", {
    "sample_selector_config": {"desired_samples_per_component": 200, "ratio_reused_samples_to_desired": 0.0},
    "model_initialization": {"num_initial_components": 45}}))
# C2: examples/6_samtron_planar4.py (D=10, 100 initial components, 100 samples per component)
timed("C2 SAMTRON planar_robot_4 (ex.6)", runner_for("planar_robot_4", "SAMTRON", {
    "num_component_adapter_config": {"del_iters": 10, "add_iters": 1},
    "sample_selector_config": {"desired_samples_per_component": 100, "ratio_reused_samples_to_desired": 0.0},
    "model_initialization": {"num_initial_components": 100}}))
