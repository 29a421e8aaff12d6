"""tcgen05 log-density kernels (3xTF32 streamed, 2 x fp16 resident) against the fp64 oracle and the exact-fp32
SIMT kernel."""
import numpy as np
import pytest
import torch

import oracle as O
from test_kernels_gpu import make_problem, gmm32_of, dev, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["tf32", "h16"])
@pytest.mark.parametrize("K,D,N", [(1, 32, 1), (3, 32, 127), (5, 64, 128), (4, 96, 129), (3, 128, 1000),
                                   (7, 256, 513), (2, 224, 300)])
def test_tc_logdens_matches_oracle(K, D, N, kind):
    from gmmvi_b200 import ops
    g, X = make_problem(K, D, N, seed=100 + D, scale=30.0)
    g32 = gmm32_of(g)
    linv, prec, cst, ok = ops.prepare_full(dev(g32.chol_cov))
    Xd, md = dev(X), dev(g32.means)
    lq_tc = ops.logdens_full(Xd, md, linv, cst, memo=False, tensor_cores=kind)
    lq_simt = ops.logdens_full(Xd, md, linv, cst, memo=False, tensor_cores=False)
    torch.cuda.synchronize()
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), False)
    ref = O.component_log_densities(g_in, X.astype(np.float64))
    assert rel_err(lq_simt.cpu().numpy(), ref) < 1e-5
    assert rel_err(lq_tc.cpu().numpy(), ref) < 1e-5
    # element-wise: the Mahalanobis part must agree to ~1e-6 of its own magnitude
    err = np.abs(lq_tc.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 5e-6, err.max()


@pytest.mark.parametrize("kind,D", [("tf32", 64), ("h16", 64), ("h16", 128), ("h16", 192), ("h16", 256)])
def test_tc_logdens_many_tiles_and_components(kind, D):
    """More work items than SMs: exercises the persistent loop, both TMEM buffers, the stage ring wrap and (h16) the
    reload of the resident factor when a CTA's work range crosses a component boundary.  D = 64 runs the kernel with the
    A operand in shared memory, D >= 128 the one with A in tensor memory (4, 6 and 8 stages per work item: the ring of
    8 A stages wraps differently for each) and its per-CTA cache of mean rows."""
    from gmmvi_b200 import ops
    K, N = 40, 128 * 12 + 5
    g, X = make_problem(K, D, N, seed=7, scale=10.0)
    g32 = gmm32_of(g)
    linv, prec, cst, ok = ops.prepare_full(dev(g32.chol_cov))
    Xd, md = dev(X), dev(g32.means)
    a = ops.logdens_full(Xd, md, linv, cst, memo=False, tensor_cores=kind)
    b = ops.logdens_full(Xd, md, linv, cst, memo=False, tensor_cores=False)
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-6
    # deterministic: same bits on a second launch
    a2 = ops.logdens_full(Xd, md, linv, cst, memo=False, tensor_cores=kind)
    assert torch.equal(a, a2)


@pytest.mark.parametrize("K,D,N", [(3, 100, 700), (2, 200, 257), (6, 20, 300), (2, 12, 64), (300, 64, 130), (2, 36, 90),
                                   (4, 160, 400), (3, 192, 260), (300, 128, 130), (6000, 128, 100)])
def test_h16_logdens_odd_dims(K, D, N):
    """fp16 path on dimensions that are not multiples of 64 (zero-padded operand, clamped loads), and with more components
    than SMs (several factor reloads per CTA).  K = 6000 with one sample tile gives every CTA more components than the
    TMEM-A kernel caches mean rows for: the entry point must fall back to the shared-memory-A kernel."""
    from gmmvi_b200 import ops
    g, X = make_problem(K, D, N, seed=300 + D, scale=30.0)
    g32 = gmm32_of(g)
    linv, prec, cst, ok = ops.prepare_full(dev(g32.chol_cov))
    Xd, md = dev(X), dev(g32.means)
    a = ops.logdens_full(Xd, md, linv, cst, memo=False, tensor_cores="h16")
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), False)
    ref = O.component_log_densities(g_in, X.astype(np.float64))
    err = np.abs(a.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 5e-6, err.max()


@pytest.mark.parametrize("mean_scale,cov_scale", [(1000.0, 1e-4), (1e-3, 1e-6), (3e4, 1.0), (0.0, 1e4)])
def test_h16_logdens_dynamic_range(mean_scale, cov_scale):
    """The power-of-two operand scaling must hold the 1e-5 tolerance when the means are far from the origin and the
    covariances tiny (x - mu cancels to 1e-5 of |x|), tiny everything, and huge covariances."""
    from gmmvi_b200 import ops
    K, D, n_per = 4, 64, 100
    rng = np.random.default_rng(11)
    means = (rng.standard_normal((K, D)) * mean_scale).astype(np.float32)
    A = rng.standard_normal((K, D, D))
    cov = (A @ A.transpose(0, 2, 1) / D + np.eye(D)) * cov_scale
    chol = np.linalg.cholesky(cov).astype(np.float32)
    X = np.concatenate([means[k] + (chol[k].astype(np.float64) @ rng.standard_normal((D, n_per))).T for k in range(K)])
    X = X.astype(np.float32)
    linv, prec, cst, ok = ops.prepare_full(dev(chol))
    a = ops.logdens_full(dev(X), dev(means), linv, cst, memo=False, tensor_cores="h16")
    g_in = O.OracleGMM(np.zeros(K), means.astype(np.float64), chol.astype(np.float64), False)
    ref = O.component_log_densities(g_in, X.astype(np.float64))
    own = np.repeat(np.arange(K), n_per)
    got = a.cpu().numpy()
    # own-component entries are the cancellation-sensitive ones
    e_own = np.abs(got[own, np.arange(K * n_per)] - ref[own, np.arange(K * n_per)]) / np.abs(ref[own, np.arange(K * n_per)])
    assert e_own.max() < 1e-5, e_own.max()
    # all entries: 1e-5 relative (north_star tolerance for log densities)
    fin = np.isfinite(ref)
    e_all = np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0)
    assert e_all.max() < 1e-5, e_all.max()


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("b,M,N,Kd", [(3, 256, 256, 256), (2, 70, 45, 132), (1, 300, 520, 64), (5, 128, 16, 36)])
def test_tc_bgemm(ta, tb, b, M, N, Kd):
    from gmmvi_b200 import ops
    rng = np.random.default_rng(70 + M)
    A = rng.standard_normal((b, Kd, M) if ta else (b, M, Kd)).astype(np.float32)
    B = rng.standard_normal((b, N, Kd) if tb else (b, Kd, N)).astype(np.float32)
    ref = (A.transpose(0, 2, 1) if ta else A).astype(np.float64) @ (B.transpose(0, 2, 1) if tb else B).astype(np.float64)
    out = ops.bgemm(dev(A), dev(B), ta, tb, 0.5, tensor_cores=True)
    simt = ops.bgemm(dev(A), dev(B), ta, tb, 0.5, tensor_cores=False)
    assert rel_err(simt.cpu().numpy(), 0.5 * ref) < 1e-5
    assert rel_err(out.cpu().numpy(), 0.5 * ref) < 5e-6, rel_err(out.cpu().numpy(), 0.5 * ref)


@pytest.mark.parametrize("K,D,N,dense", [(2, 256, 5000, True), (3, 100, 4500, True), (150, 64, 700, True),
                                          (5, 48, 3000, False), (2, 200, 2600, False)])
def test_tc_stein_stats(K, D, N, dense):
    """tcgen05 Stein statistics against an fp64 einsum: several drains of the TMEM accumulator per component
    (N > 2048 per work unit), split block ranges (K < #SMs), padded D, and skipped blocks (`dense=False`)."""
    from gmmvi_b200 import ops
    rng = np.random.default_rng(900 + D)
    X = (rng.standard_normal((N, D)) * 3 + rng.standard_normal(D) * 20).astype(np.float32)
    means = (rng.standard_normal((K, D)) * 20).astype(np.float32)
    G = rng.standard_normal((N, D)).astype(np.float32)
    W = rng.random((K, N)).astype(np.float32) ** 4
    act = np.ones((K, (N + 127) // 128), np.uint8)
    if not dense:          # zero the weight of whole 128-sample blocks and mark them inactive
        for k in range(K):
            off = rng.random(act.shape[1]) < 0.6
            off[k % act.shape[1]] = False
            act[k, off] = 0
            for b in np.nonzero(off)[0]:
                W[k, b * 128:(b + 1) * 128] = 0.0
    W = (W / W.sum(1, keepdims=True)).astype(np.float32)
    A = rng.standard_normal((K, D, D))
    prec = (A @ A.transpose(0, 2, 1) / D + np.eye(D)).astype(np.float32)
    Hneg, gneg = ops.stein_full(dev(X), dev(means), dev(prec), dev(W), dev(act, torch.uint8), dev(G), True)
    X64, W64, G64 = X.astype(np.float64), W.astype(np.float64), G.astype(np.float64)
    M = np.einsum("kn,knj,ni->kji", W64, X64[None] - means.astype(np.float64)[:, None], G64)
    T = prec.astype(np.float64) @ M
    Href = -0.5 * (T + T.transpose(0, 2, 1))
    gref = -(W64 @ G64)
    assert rel_err(Hneg.cpu().numpy(), Href) < 1e-4, rel_err(Hneg.cpu().numpy(), Href)
    assert rel_err(gneg.cpu().numpy(), gref) < 1e-4
    # bitwise reproducible
    H2, _ = ops.stein_full(dev(X), dev(means), dev(prec), dev(W), dev(act, torch.uint8), dev(G), True)
    assert torch.equal(Hneg, H2)


def test_full_size_c5_properties():
    """BASELINE's stress configuration at full size (K = 512, D = 256, N = 65536): the oracle cannot run it, so the tensor-core
    pass is checked (a) against the exact-fp32 SIMT kernel on a slice of the components, (b) for exact scaling
    invariance (x, mu, L -> 4 x, 4 mu, 4 L shifts every density by -D log 4 and nothing else), (c) through the
    chi-square law of the samples' own components, (d) the importance weights of every component sum to one."""
    from gmmvi_b200 import ops
    K, D, N = 512, 256, 65536
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((K, D, D), device="cuda", generator=g)
    chol = torch.linalg.cholesky(A @ A.transpose(1, 2) / D + torch.eye(D, device="cuda")).contiguous()
    del A
    means = (torch.randn((K, D), device="cuda", generator=g) * 31.63).contiguous()
    comp = torch.arange(N, device="cuda") // (N // K)
    eps = torch.randn((N, D), device="cuda", generator=g)
    X = (means[comp] + torch.einsum("nij,nj->ni", chol[comp], eps)).contiguous()
    linv, prec, cst, ok = ops.prepare_full(chol)
    assert bool(ok.all())
    lq = ops.logdens_full(X, means, linv, cst, memo=False)
    assert lq.shape == (K, N) and bool(torch.isfinite(lq).all())
    # (a) exact fp32 kernel on 6 components spread over the range (all samples)
    sel = torch.tensor([0, 1, 100, 255, 256, 511], device="cuda")
    ref = ops.logdens_full(X, means[sel].contiguous(), linv[sel].contiguous(), cst[sel].contiguous(), memo=False,
                           tensor_cores=False)
    err = ((lq[sel] - ref).abs() / ref.abs().clamp_min(1.0)).max().item()
    assert err < 5e-6, err
    # (b) scaling by a power of two is exact in fp32: x, mu, L -> 4 x, 4 mu, 4 L changes every density by -D log 4 only
    lq4 = ops.logdens_full((4.0 * X).contiguous(), (4.0 * means).contiguous(), (0.25 * linv).contiguous(),
                           (cst - D * float(np.log(4.0))).contiguous(), memo=False)
    err = ((lq4 + D * float(np.log(4.0)) - lq).abs() / lq.abs().clamp_min(1.0)).max().item()
    assert err < 1e-6, err
    # (c) own component: -2 (lq - cst) = |eps|^2 is chi-square with D degrees of freedom
    own = lq[comp, torch.arange(N, device="cuda")]
    maha = -2.0 * (own - cst[comp])
    assert abs(maha.mean().item() - D) < 0.5 and abs(maha.var().item() - 2 * D) < 0.05 * 2 * D
    assert ((maha - (eps * eps).sum(1)).abs() / D).max().item() < 1e-3
    # (d) self-normalised importance weights
    bg = ops.mixture_lse(lq, torch.full((K,), -float(np.log(K)), device="cuda"))
    W = ops.importance_weights(lq, bg, None, True, None, True, False, False, False)["W"]
    assert (W.sum(1) - 1.0).abs().max().item() < 1e-5 and bool((W >= 0).all())


@pytest.mark.parametrize("K,D,N,scale", [(3, 128, 300, 30.0), (5, 256, 513, 30.0), (6, 100, 700, 0.3), (40, 256, 1300, 0.2),
                                         (4, 224, 129, 5.0), (70, 72, 400, 0.1)])
def test_tc_mixture_gradient(K, D, N, scale):
    """Tensor-core mixture gradient (mg::mixgrad_h16_kernel) against the fp64 oracle and the SIMT tile engine.  Small
    `scale` = overlapping components: every tile carries many (tile, component) items whose contributions are summed
    by the read-modify-write epilogue; K = 70 spans three mask words; N is ragged; D = 100 / 224 / 72 are zero padded."""
    from gmmvi_b200 import ops
    g, X = make_problem(K, D, N, seed=400 + D + K, scale=scale)
    g32 = gmm32_of(g)
    linv, prec, cst, ok = ops.prepare_full(dev(g32.chol_cov))
    Xd, md, lw = dev(X), dev(g32.means), dev(g32.log_weights)
    lq = ops.logdens_full(Xd, md, linv, cst, memo=False)
    logq = ops.mixture_lse(lq, lw)
    g_tc = ops.mixture_grad_full(Xd, md, prec, lq, lw, logq, tensor_cores=True)
    g_simt = ops.mixture_grad_full(Xd, md, prec, lq, lw, logq, tensor_cores=False)
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), False)
    _, ref, _ = O.log_density_and_grad(g_in, X.astype(np.float64))
    assert rel_err(g_simt.cpu().numpy(), ref) < 1e-4
    assert rel_err(g_tc.cpu().numpy(), ref) < 1e-4
    assert rel_err(g_tc.cpu().numpy(), g_simt.cpu().numpy()) < 2e-5
    # deterministic: the same bits on a second launch
    assert torch.equal(g_tc, ops.mixture_grad_full(Xd, md, prec, lq, lw, logq, tensor_cores=True))
