"""torch.profiler view (host and device time) of the C1 example configuration (examples/5_samtron_20D_student-T.py),
a small problem whose iteration is bound by launches and host glue.  usage: python profiles/prof_config.py [iters]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]] + sys.argv[1:]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
from gmmvi_b200.configs import get_default_algorithm_config, get_default_experiment_config, update_config  # noqa: E402
from gmmvi_b200.gmmvi_runner import GmmviRunner  # noqa: E402

algo = update_config(get_default_algorithm_config("SAMTRON"), {
    "sample_selector_config": {"desired_samples_per_component": 200, "ratio_reused_samples_to_desired": 0.0},
    "model_initialization": {"num_initial_components": 45}})
config = update_config(update_config(get_default_experiment_config("stm20"), {"start_seed": 1}), algo)
config["gmmvi_runner_config"] = {"log_metrics_interval": 10 ** 9}
runner = GmmviRunner.build_from_config(config)
for n in range(15):
    runner.iterate_and_log(n)
torch.cuda.synchronize()
t0 = time.perf_counter()
for n in range(15, 15 + iters):
    runner.iterate_and_log(n)
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / iters * 1e3:.2f} ms / iteration without the profiler")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for n in range(15 + iters, 15 + 2 * iters):
        runner.iterate_and_log(n)
    torch.cuda.synchronize()
ev = prof.key_averages()
print("--- host (self CPU time per iteration, ms) ---")
for e in sorted(ev, key=lambda e: -e.self_cpu_time_total)[:25]:
    print(f"{e.self_cpu_time_total / iters / 1e3:8.3f}  n={e.count / iters:6.1f}  {e.key[:90]}")
print("--- device (self device time per iteration, ms) ---")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:15]:
    print(f"{e.self_device_time_total / iters / 1e3:8.3f}  n={e.count / iters:6.1f}  {e.key[:90]}")
