"""Small driver used for the ncu captures under profiles/: runs one hot kernel a few times at the C5 size.
usage: python profiles/prof_kernels.py {logdens|update|prepare|stein|grad} [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmmvi_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "logdens"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
K, D, N = 512, 256, 65536
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn((K, D, D), device="cuda", generator=g)
cov = A @ A.transpose(1, 2) / D + torch.eye(D, device="cuda")
chol = torch.linalg.cholesky(cov).contiguous()
means = (torch.randn((K, D), device="cuda", generator=g) * 31.63).contiguous()
comp = torch.arange(N, device="cuda") // (N // K)
X = (means[comp] + torch.einsum("nij,nj->ni", chol[comp], torch.randn((N, D), device="cuda", generator=g))).contiguous()
linv, prec, cst, ok = ops.prepare_full(chol)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    print(f"{which}: {ev0.elapsed_time(ev1) / reps:.3f} ms per launch")


if which == "logdens":
    timed(lambda: ops.logdens_full(X, means, linv, cst, memo=False))
elif which == "logdens_tf32":
    timed(lambda: ops.logdens_full(X, means, linv, cst, memo=False, tensor_cores="tf32"))
elif which == "logdens_simt":
    timed(lambda: ops.logdens_full(X, means, linv, cst, memo=False, tensor_cores=False))
elif which in ("stein", "stein_simt"):
    if which == "stein_simt":
        os.environ["GMMVI_B200_TC_STEIN"] = "0"
    # dense case: every sample carries weight for every component
    Nd = 16384
    Wd = torch.rand((K, Nd), device="cuda", generator=g)
    Wd = (Wd / Wd.sum(1, keepdim=True)).contiguous()
    act = torch.ones((K, Nd // 128), device="cuda", dtype=torch.uint8)
    Gd = torch.randn((Nd, D), device="cuda", generator=g).contiguous()
    Xd = X[:Nd].contiguous()
    timed(lambda: ops.stein_full(Xd, means, prec, Wd, act, Gd, True))
    print("dense Stein: K=%d D=%d N=%d -> %.2f TFLOP algorithmic (2 D^2 per pair)" % (K, D, Nd, 2.0 * D * D * K * Nd / 1e12))
elif which == "prepare":
    timed(lambda: ops.prepare_full(chol))
elif which == "update":
    H = (A @ A.transpose(1, 2) / D * 0.3).contiguous()
    gn = torch.randn((K, D), device="cuda", generator=g).contiguous()
    steps = torch.full((K,), 0.1, device="cuda")
    last = torch.full((K,), 30.0, device="cuda")
    timed(lambda: ops.update_components("trust-region", False, means, chol, H, gn, steps, last, None, 1.0))
    print("evals/component", ops.last_update_evals.float().mean().item())
