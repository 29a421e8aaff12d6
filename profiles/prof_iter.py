"""torch.profiler view of a few SAMTRON iterations on the bench workload (rank 0 prints).  Run under torchrun for N > 1:
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/prof_iter.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gmmvi_b200 import rng  # noqa: E402
from gmmvi_b200.distributed import ShardContext  # noqa: E402
from gmmvi_b200.experiments.target_distributions.gmm import GMM_LNPDF  # noqa: E402
from gmmvi_b200.models.full_cov_gmm import FullCovGMM  # noqa: E402
from gmmvi_b200.models.gmm_wrapper import GmmWrapper  # noqa: E402
from gmmvi_b200.optimization.gmmvi import GMMVI  # noqa: E402

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
K, D, per = bench.K_COMP, bench.DIM, bench.PER_COMP
# --dense: overlapping components (mean scale 0.05): no (component, sample block) can be skipped
means, chols, tmeans, tchols = bench.workload_arrays(K, D, 0, 0.05 if "--dense" in sys.argv else bench.PRIOR_SCALE)
model = FullCovGMM.from_cholesky(np.ones(K, np.float32) / K, means, chols, device=dev)
tgt = GMM_LNPDF.from_cholesky(np.ones(10) / 10, tmeans, tchols, device=dev)
cfg = bench.samtron_config(per)
g = GMMVI.build_from_config(cfg, tgt, GmmWrapper.build_from_config(model, cfg))
if world > 1:
    g.enable_sharding(ShardContext(rank, world))
rng.set_seed(1234)
if "--graph" in sys.argv:
    g.enable_cuda_graph()
for _ in range(4):
    g.train_iter()
torch.cuda.synchronize()
steps = 5
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.train_iter()
    e1.record()
    torch.cuda.synchronize()
if rank == 0:
    print(f"wall per iteration (profiler on): {e0.elapsed_time(e1) / steps:.3f} ms")
    ev = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.self_device_time_total for e in ev)
    print(f"sum of device time per iteration: {tot / steps / 1e3:.3f} ms")
    for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:70]:
        print(f"{e.self_device_time_total / steps / 1e3:8.3f} ms  n={e.count / steps:5.1f}  {e.key[:90]}")
bench.shutdown_distributed([g])
