"""Debug aid: does the MORE iteration (N not a multiple of 64) read uninitialised memory?  Eager run vs eager run whose
allocator state is perturbed with NaN-filled garbage between iterations."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gmmvi_b200 import rng  # noqa: E402
from test_graph_gpu import _fixed  # noqa: E402


def run(perturb, iters=8, N=900, variant="lin-more"):
    rng.set_seed(11)
    torch.manual_seed(123)
    g = _fixed(3, 6, N, "trust-region", False, variant)
    out = []
    for i in range(iters):
        if perturb:
            junk = [torch.full((s,), float("nan") if perturb == "nan" else 1e30, device="cuda") for s in (1 << 10, 1 << 14, 1 << 18, 1 << 20, 1 << 22, 3 << 20)]
            del junk
        g.train_iter()
        out.append(g.model.means.detach().cpu().numpy().copy())
    return out


for route in ("1", "tf32", "0"):
    os.environ["GMMVI_B200_MORE_TC"] = route
    a, b, c = run(None), run("nan"), run("big")
    print("route", route, "clean vs NaN-garbage:", [float(np.nanmax(np.abs(x - y))) if np.isfinite(y).all() else "nan" for x, y in zip(a, b)])
    print("route", route, "clean vs 1e30-garbage:", [float(np.max(np.abs(x - y))) if np.isfinite(y).all() else "nan" for x, y in zip(a, c)])
