"""Debug aid: is the MORE iteration with the mixture-based selector reproducible run to run (eager), and where do eager and
graph runs start to differ?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gmmvi_b200 import rng  # noqa: E402
from test_graph_gpu import _fixed  # noqa: E402


def run(graph, iters, variant="lin-more", N=900):
    rng.set_seed(11)
    torch.manual_seed(123)
    g = _fixed(3, 6, N, "trust-region", False, variant)
    if graph:
        g.enable_cuda_graph()
    out = []
    for _ in range(iters):
        g.train_iter()
        out.append(g.model.means.detach().cpu().numpy().copy())
    return out


for variant, N in (("lin-more", 900), ("lin-more", 1024), ("more", 300)):
    a, b, c = run(False, 8, variant, N), run(False, 8, variant, N), run(True, 8, variant, N)
    print(variant, N, "eager vs eager:", [float(np.max(np.abs(x - y))) for x, y in zip(a, b)])
    print(variant, N, "eager vs graph:", [float(np.max(np.abs(x - y))) for x, y in zip(a, c)])
