"""Accuracy / speed of the tensor-core Stein statistics against the accumulator drain interval (GMMVI_B200_STEIN_FLUSH):
dense weights, D = 256; error of H against an fp64 evaluation of the same sums on the device."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gmmvi_b200 import ops  # noqa: E402

K, D, N = 16, 256, 65536
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn((K, D, D), device="cuda", generator=g)
chol = torch.linalg.cholesky(A @ A.transpose(1, 2) / D + torch.eye(D, device="cuda")).contiguous()
means = (torch.randn((K, D), device="cuda", generator=g) * 0.05).contiguous()
X = (torch.randn((N, D), device="cuda", generator=g)).contiguous()
W = torch.rand((K, N), device="cuda", generator=g)
W = (W / W.sum(1, keepdim=True)).contiguous()
G = torch.randn((N, D), device="cuda", generator=g).contiguous()
act = torch.ones((K, N // 128), device="cuda", dtype=torch.uint8)
linv, prec, cst, _ = ops.prepare_full(chol)
Xc = X.double().unsqueeze(0) - means.double().unsqueeze(1)                      # [K, N, D]
M = torch.einsum("kn,knj,ni->kji", W.double(), Xc, G.double())
T = prec.double() @ M
Href = -0.5 * (T + T.transpose(1, 2))
for flush in (8, 16, 32, 64, 128, 512):
    os.environ["GMMVI_B200_STEIN_FLUSH"] = str(flush)
    H, gn = ops.stein_full(X, means, prec, W, act, G, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.stein_full(X, means, prec, W, act, G, True)
    e1.record()
    torch.cuda.synchronize()
    err = float((H.double() - Href).abs().max() / Href.abs().max())
    print(f"flush {flush:4d}: rel err {err:.3e}   {e0.elapsed_time(e1) / 3:.3f} ms per call (K={K})")
