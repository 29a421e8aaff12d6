"""Debug aid: eager and graph runs of the MORE + mixture-based-selector iteration in lockstep; prints, per iteration, which
intermediate first differs (samples, target values, NG estimate, update outputs, weights)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gmmvi_b200 import rng  # noqa: E402
from test_graph_gpu import _fixed  # noqa: E402


def make(graph):
    rng.set_seed(11)
    torch.manual_seed(123)
    g = _fixed(3, 6, 900, "trust-region", False, "lin-more")
    est = g.ng_estimator
    orig = est.get_expected_hessian_and_grad

    def patched(*a):
        H, gr = orig(*a)
        est.last_H, est.last_g = H, gr
        return H, gr
    est.get_expected_hessian_and_grad = patched
    if graph:
        g.enable_cuda_graph()
    return g


def snap(g):
    m = g.model
    d = {"samples": g.sample_db.samples, "lnpdfs": g.sample_db.target_lnpdfs, "H": g.ng_estimator.last_H, "g": g.ng_estimator.last_g,
         "ok": g.ng_estimator.last_ok.float(), "etas": g.ng_based_updater.last_etas, "kls": g.ng_based_updater.last_kls,
         "succ": g.ng_based_updater.last_success.float(), "means": m.means, "chol": m.chol_cov, "logw": m.log_weights,
         "l2": m.l2_regularizers, "steps": m.stepsizes, "rewards": m.reward_history}
    return {k: v.detach().cpu().numpy().copy() for k, v in d.items()}


# torch's generator is global: run the two in sequence, saving the eager snapshots first
e = make(False)
es = []
for i in range(7):
    e.train_iter()
    es.append(snap(e))
gr = make(True)
for i in range(7):
    gr.train_iter()
    torch.cuda.synchronize()
    s = snap(gr)
    diff = {k: float(np.max(np.abs(s[k] - es[i][k]))) for k in s if s[k].shape == es[i][k].shape}
    print(i, {k: v for k, v in diff.items() if v != 0.0} or "identical")
