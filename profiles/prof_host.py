"""Host-side (Python) cost of one iteration of the launch-bound BASELINE configurations C1 / C2: cProfile over the timed
iterations + launches per iteration.  usage: python profiles/prof_host.py [C1|C2] [iterations]"""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmmvi_b200 import ops  # noqa: E402
from gmmvi_b200.configs import get_default_algorithm_config, get_default_experiment_config, update_config  # noqa: E402
from gmmvi_b200.gmmvi_runner import GmmviRunner  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "C2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
if which == "C1":
    exp, over = "stm20", {"sample_selector_config": {"desired_samples_per_component": 200, "ratio_reused_samples_to_desired": 0.0},
                          "model_initialization": {"num_initial_components": 45}}
else:
    exp, over = "planar_robot_4", {"num_component_adapter_config": {"del_iters": 10, "add_iters": 1},
                                   "sample_selector_config": {"desired_samples_per_component": 100, "ratio_reused_samples_to_desired": 0.0},
                                   "model_initialization": {"num_initial_components": 100}}
algo = update_config(get_default_algorithm_config("SAMTRON"), over)
config = update_config(update_config(get_default_experiment_config(exp), {"start_seed": 1}), algo)
config["gmmvi_runner_config"] = {"log_metrics_interval": 10 ** 9}
runner = GmmviRunner.build_from_config(config)
for n in range(15):
    runner.iterate_and_log(n)
torch.cuda.synchronize()
k0, c0 = ops.KERNELS, ops.LAUNCHES
t0 = time.perf_counter()
for n in range(15, 15 + iters):
    runner.iterate_and_log(n)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / iters
print(f"{which}: {dt * 1e3:.2f} ms/iter, {(ops.KERNELS - k0) / iters:.0f} gvi kernels and {(ops.LAUNCHES - c0) / iters:.0f} "
      f"C-ABI calls per iteration, K={runner.gmmvi.model.num_components}")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for n in range(15 + iters, 15 + iters + 5):
        runner.iterate_and_log(n)
    torch.cuda.synchronize()
ev = prof.key_averages()
dev_ms = sum(e.self_device_time_total for e in ev) / 5 / 1e3
n_launch = sum(e.count for e in ev if e.key in ("cudaLaunchKernel", "cudaLaunchKernelExC", "cuLaunchKernel", "cuLaunchKernelEx")) / 5
print(f"device time {dev_ms:.3f} ms/iter, {n_launch:.0f} kernel launches (all, incl. torch) per iteration")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:14]:
    print(f"    {e.self_device_time_total / 5 / 1e3:8.3f} ms  n={e.count / 5:6.1f}  {e.key[:90]}")
pr = cProfile.Profile()
pr.enable()
for n in range(20 + iters, 20 + 2 * iters):
    runner.iterate_and_log(n)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print("\n".join(l[:170] for l in s.getvalue().splitlines()[:80]))
