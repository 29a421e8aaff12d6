"""Debug aid: the DEFERRED selection path run eagerly (no graph) against the standard eager path, MORE + mixture-based
selector, N = 900: separates "the deferred path computes something else" from "the graph replay computes something else"."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gmmvi_b200 import rng  # noqa: E402
from test_graph_gpu import _fixed  # noqa: E402


def run(mode, iters=7):
    rng.set_seed(11)
    torch.manual_seed(123)
    g = _fixed(3, 6, 900, "trust-region", False, "lin-more")
    out = []
    for i in range(iters):
        if mode == "std":
            s = g.sample_selector.select_samples()
        else:
            s, payload = g.sample_selector.select_samples_deferred()
            g.sample_db.add_samples(s[0], g.model.means, g.model.chol_cov, s[3], s[4], s[1], prepared=g.sample_selector._prepared())
        g._run_updates(*s)
        out.append({"bg": s[2].detach().cpu().numpy().copy(), "mapping": s[1].detach().cpu().numpy().copy(),
                    "means": g.model.means.detach().cpu().numpy().copy()})
    return out


a, b = run("std"), run("deferred")
for i, (x, y) in enumerate(zip(a, b)):
    print(i, {k: float(np.max(np.abs(x[k].astype(np.float64) - y[k]))) for k in x})
