"""Small driver used for the ncu captures under profiles/: runs one hot kernel a few times at the C5 size.
usage: python profiles/prof_kernels.py {logdens|update|prepare|stein|stein_full|mixgrad|gsum|small|diag} [reps]"""
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
elif which in ("stein_full", "mixgrad", "gsum"):
    # the C5 shape with EVERY (component, sample block) active: overlapping components (bench.py's dense variant)
    means_d = (torch.randn((K, D), device="cuda", generator=g) * 0.05).contiguous()
    Xd = (means_d[comp] + torch.einsum("nij,nj->ni", chol[comp], torch.randn((N, D), device="cuda", generator=g))).contiguous()
    lq = ops.logdens_full(Xd, means_d, linv, cst, memo=False)
    logw = torch.full((K,), -float(np.log(K)), device="cuda")
    logq = ops.mixture_lse(lq, logw)
    if which == "mixgrad":
        timed(lambda: ops.mixture_grad_full(Xd, means_d, prec, lq, logw, logq))
        print("dense mixture gradient: %.2f TFLOP algorithmic (2 D^2 per pair)" % (2.0 * D * D * K * N / 1e12))
    else:
        iw = ops.importance_weights(lq, logq, None, True, None, True, False, False, True)
        Gd = torch.randn((N, D), device="cuda", generator=g).contiguous()
        print("active blocks: %.3f" % iw["active"].float().mean().item())
        timed(lambda: ops.stein_full(Xd, means_d, prec, iw["W"], iw["active"], Gd, True))
        print("dense Stein at full C5 size: K=%d D=%d N=%d -> %.2f TFLOP algorithmic (2 D^2 per pair)" % (K, D, N, 2.0 * D * D * K * N / 1e12))
elif which == "small":
    # BASELINE C2 shape: D = 10, K = 128, N = 12800 (the small-dimension kernels of csrc/small_dim.cu)
    Ks, Ds, Ns = 128, 10, 12800
    As = torch.randn((Ks, Ds, Ds), device="cuda", generator=g)
    chs = torch.linalg.cholesky(As @ As.transpose(1, 2) / Ds + torch.eye(Ds, device="cuda")).contiguous()
    ms = torch.randn((Ks, Ds), device="cuda", generator=g).contiguous()
    Xs = (ms[torch.arange(Ns, device="cuda") // 100] + torch.randn((Ns, Ds), device="cuda", generator=g)).contiguous()
    li, pr, cs, _ = ops.prepare_full(chs)
    timed(lambda: ops.logdens_full(Xs, ms, li, cs, memo=False))
    lq = ops.logdens_full(Xs, ms, li, cs, memo=False)
    logw = torch.full((Ks,), -float(np.log(Ks)), device="cuda")
    logq = ops.mixture_lse(lq, logw)
    timed(lambda: ops.mixture_grad_full(Xs, ms, pr, lq, logw, logq))
    iw = ops.importance_weights(lq, logq, None, True, None, True, False, False, True)
    Gs = torch.randn((Ns, Ds), device="cuda", generator=g).contiguous()
    timed(lambda: ops.stein_full(Xs, ms, pr, iw["W"], iw["active"], Gs, True))
elif which == "diag":
    # BASELINE C4-diagonal shape: D = 200, K = 256, N = 16384 (csrc/diag.cu)
    Kd, Dd, Nd = 256, 200, 16384
    md = (torch.randn((Kd, Dd), device="cuda", generator=g) * 100).contiguous()
    sd = (torch.rand((Kd, Dd), device="cuda", generator=g) * 10 + 10).contiguous()
    Xd = (md[torch.arange(Nd, device="cuda") // 64] + sd[torch.arange(Nd, device="cuda") // 64] * torch.randn((Nd, Dd), device="cuda", generator=g)).contiguous()
    timed(lambda: ops.logdens_diag(Xd, md, sd))
    lq = ops.logdens_diag(Xd, md, sd)
    logw = torch.full((Kd,), -float(np.log(Kd)), device="cuda")
    logq = ops.mixture_lse(lq, logw)
    iw = ops.importance_weights(lq, logq, None, True, None, True)
    Gd = torch.randn((Nd, Dd), device="cuda", generator=g).contiguous()
    timed(lambda: ops.stein_diag(Xd, md, sd, iw["W"], Gd))
elif which == "prepare":
    timed(lambda: ops.prepare_full(chol))
elif which == "update":
    H = (A @ A.transpose(1, 2) / D * 0.3).contiguous()
    gn = torch.randn((K, D), device="cuda", generator=g).contiguous()
    steps = torch.full((K,), 0.1, device="cuda")
    last = torch.full((K,), 30.0, device="cuda")
    timed(lambda: ops.update_components("trust-region", False, means, chol, H, gn, steps, last, None, 1.0))
    print("evals/component", ops.last_update_evals.float().mean().item())
