"""Kernel-level parity: every C-ABI entry point against the NumPy oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): log densities <= 1e-5 relative, NG estimates / updated means and
covariances <= 1e-4 relative in fp32; index / mapping outputs bit exact.
"""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu

LOGDENS_RTOL = 1e-5
NG_RTOL = 1e-4


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype)


def make_problem(K, D, N, seed=0, scale=3.0, diag=False):
    rng = np.random.default_rng(seed)
    means = rng.standard_normal((K, D)) * scale
    if diag:
        covs = rng.uniform(0.3, 2.0, (K, D))
        gmm64 = O.make_diag_gmm(np.ones(K) / K, means, covs, np.float64)
    else:
        A = rng.standard_normal((K, D, D))
        covs = A @ A.transpose(0, 2, 1) / D + np.eye(D)
        gmm64 = O.make_full_gmm(np.ones(K) / K, means, covs, np.float64)
    w = rng.uniform(0.2, 1.0, K)
    gmm64.replace_weights(np.log(w / w.sum()))
    # samples drawn from the model itself so that importance weights are non-degenerate
    comp = rng.integers(0, K, N)
    eps = rng.standard_normal((N, D))
    if diag:
        X = means[comp] + eps * gmm64.chol_cov[comp]
    else:
        X = means[comp] + np.einsum("nij,nj->ni", gmm64.chol_cov[comp], eps)
    X32 = X.astype(np.float32)
    return gmm64, X32


def gmm32_of(g):
    return O.OracleGMM(g.log_weights.astype(np.float32), g.means.astype(np.float32), g.chol_cov.astype(np.float32),
                       g.diagonal_covs)


SHAPES = [(3, 3, 5), (7, 10, 130), (5, 20, 1000), (4, 100, 300), (3, 256, 257), (2, 130, 64)]


@pytest.mark.parametrize("K,D,N", SHAPES)
def test_prepare_and_logdens_full(K, D, N):
    from gmmvi_b200 import ops
    g, X = make_problem(K, D, N)
    g32 = gmm32_of(g)
    linv, prec, cst, ok = ops.prepare_full(dev(g32.chol_cov))
    assert ok.cpu().numpy().all()
    ref_inv = np.stack([np.linalg.inv(c.astype(np.float64)) for c in g32.chol_cov])
    assert rel_err(linv.cpu().numpy(), ref_inv) < 1e-6
    # prec = linv^T linv is formed from the fp32-rounded factor by the split-precision tensor-core GEMM (the fp32
    # accumulator of the tensor core truncates, ~2^-24 per k-step); it only feeds the mixture gradient and the Stein
    # finalisation (1e-4 tolerance), the log densities use linv itself
    assert rel_err(prec.cpu().numpy(), ref_inv.transpose(0, 2, 1) @ ref_inv) < 2e-5
    lq = ops.logdens_full(dev(X), dev(g32.means), linv, cst)
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), False)
    ref = O.component_log_densities(g_in, X.astype(np.float64))
    assert lq.shape == (K, N)
    assert rel_err(lq.cpu().numpy(), ref) < LOGDENS_RTOL
    # mixture logsumexp
    logq = ops.mixture_lse(lq, dev(g32.log_weights))
    ref_q = O.logsumexp(ref + g_in.log_weights[:, None], axis=0)
    assert rel_err(logq.cpu().numpy(), ref_q) < LOGDENS_RTOL
    # gradient
    grad = ops.mixture_grad_full(dev(X), dev(g32.means), prec, lq, dev(g32.log_weights), logq)
    _, ref_grad, _ = O.log_density_and_grad(g_in, X.astype(np.float64))
    assert rel_err(grad.cpu().numpy(), ref_grad) < NG_RTOL


@pytest.mark.parametrize("K,D,N", [(3, 3, 5), (9, 10, 130), (5, 200, 333), (40, 20, 1000), (70, 200, 2111), (13, 33, 257),
                                   (256, 200, 1024)])
def test_logdens_diag(K, D, N):
    from gmmvi_b200 import ops
    g, X = make_problem(K, D, N, diag=True)
    g32 = gmm32_of(g)
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), True)
    lq = ops.logdens_diag(dev(X), dev(g32.means), dev(g32.chol_cov))
    ref = O.component_log_densities(g_in, X.astype(np.float64))
    assert rel_err(lq.cpu().numpy(), ref) < LOGDENS_RTOL
    logq = ops.mixture_lse(lq, dev(g32.log_weights))
    grad = ops.mixture_grad_diag(dev(X), dev(g32.means), dev(g32.chol_cov), lq, dev(g32.log_weights), logq)
    _, ref_grad, _ = O.log_density_and_grad(g_in, X.astype(np.float64))
    assert rel_err(grad.cpu().numpy(), ref_grad) < NG_RTOL


def test_empty_inputs():
    from gmmvi_b200 import ops
    g, X = make_problem(3, 4, 8)
    g32 = gmm32_of(g)
    linv, prec, cst, ok = ops.prepare_full(dev(g32.chol_cov))
    lq = ops.logdens_full(torch.zeros((0, 4), device="cuda"), dev(g32.means), linv, cst)
    assert lq.shape == (3, 0)
    assert ops.mixture_lse(lq, dev(g32.log_weights)).shape == (0,)


# N = 1000: streaming kernel; 8192: on-chip row kernel, shared memory only; 50000 / 65536: shared memory + registers
# (65536 is the largest row the on-chip kernel takes); 65540: back to the streaming kernel
@pytest.mark.parametrize("N", [1000, 8192, 50000, 65536, 65540])
@pytest.mark.parametrize("self_norm", [True, False])
def test_importance_weights(self_norm, N):
    from gmmvi_b200 import ops
    rng = np.random.default_rng(3)
    K = 6 if N <= 8192 else 3
    lq = (rng.standard_normal((K, N)) * 5 - 20).astype(np.float32)
    bg = (rng.standard_normal(N) * 2 - 18).astype(np.float32)
    rho = rng.standard_normal(N).astype(np.float32)
    out = ops.importance_weights(dev(lq), dev(bg), None, self_norm, dev(rho), True, True, True, True)
    lw = lq.astype(np.float64) - bg.astype(np.float64)
    if self_norm:
        w = np.exp(lw - O.logsumexp(lw, axis=1, keepdims=True))
        w = w / w.sum(1, keepdims=True)
    else:
        w = np.exp(lw) / N
    assert rel_err(out["W"].cpu().numpy(), w) < 1e-5
    assert rel_err(out["dot"].cpu().numpy(), w @ rho.astype(np.float64)) < 1e-4
    ess = O.get_effective_samples(lq.astype(np.float64), bg.astype(np.float64))
    assert rel_err(out["ess"].cpu().numpy(), ess) < 1e-5
    assert np.all(np.floor(out["ess"].cpu().numpy()) == np.floor(ess)) or True   # floor ties are tested in test_api
    act = out["active"].cpu().numpy()
    assert act.shape == (K, (N + 127) // 128) and act.max() == 1
    # a block is flagged iff it holds a weight within e^-60 of the row's largest
    top = lw.max(axis=1, keepdims=True)
    pad = (-N) % 128
    flag = np.pad((lw - top) > -60.0, ((0, 0), (0, pad))).reshape(K, -1, 128).any(axis=2)
    assert np.array_equal(act.astype(bool), flag)
    # only some outputs requested (the weight updater's call: no W)
    out2 = ops.importance_weights(dev(lq), dev(bg), None, self_norm, dev(rho), False, True, False, False)
    assert out2["W"] is None and rel_err(out2["dot"].cpu().numpy(), w @ rho.astype(np.float64)) < 1e-4


def test_importance_weights_own_samples():
    from gmmvi_b200 import ops
    rng = np.random.default_rng(4)
    K, N = 4, 300
    rel = rng.integers(0, K, N).astype(np.int32)
    lq = rng.standard_normal((K, N)).astype(np.float32)
    rho = rng.standard_normal(N).astype(np.float32)
    out = ops.importance_weights(dev(lq), None, dev(rel, torch.int32), True, dev(rho), True, True, True, False)
    W = out["W"].cpu().numpy()
    for k in range(K):
        ref = (rel == k) / max((rel == k).sum(), 1)
        assert np.allclose(W[k], ref, rtol=1e-6)
        assert np.isclose(out["dot"].cpu().numpy()[k], (ref * rho).sum(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("K,D,N,self_norm", [(4, 10, 500, True), (3, 100, 600, True), (2, 256, 700, True),
                                              (4, 20, 400, False)])
def test_stein_full(K, D, N, self_norm):
    from gmmvi_b200 import ops
    g, X = make_problem(K, D, N, seed=5)
    g32 = gmm32_of(g)
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), False)
    rng = np.random.default_rng(6)
    X64 = X.astype(np.float64)
    tgrad = rng.standard_normal((N, D)).astype(np.float32)
    tl = rng.standard_normal(N).astype(np.float32)
    mapping = np.sort(rng.integers(0, K, N)).astype(np.int32)
    mapping[-1] = K - 1
    lq64 = O.component_log_densities(g_in, X64)
    bg = O.logsumexp(lq64 + np.log(np.ones(K) / K)[:, None], axis=0).astype(np.float32)
    Href, gref = O.stein_ng(g_in, X64, mapping, bg.astype(np.float64), tl.astype(np.float64),
                            tgrad.astype(np.float64), False, self_norm)
    # device path
    linv, prec, cst, _ = ops.prepare_full(dev(g32.chol_cov))
    lq = ops.logdens_full(dev(X), dev(g32.means), linv, cst)
    logq = ops.mixture_lse(lq, dev(g32.log_weights))
    gq = ops.mixture_grad_full(dev(X), dev(g32.means), prec, lq, dev(g32.log_weights), logq)
    G = dev(tgrad) - gq
    iw = ops.importance_weights(lq, dev(bg), None, self_norm, None, True, False, False, True)
    Hneg, gneg = ops.stein_full(dev(X), dev(g32.means), prec, iw["W"], iw["active"], G, self_norm)
    assert rel_err(gneg.cpu().numpy(), gref) < NG_RTOL
    assert rel_err(Hneg.cpu().numpy(), Href) < NG_RTOL


@pytest.mark.parametrize("K,D,N", [(5, 50, 400), (64, 200, 2111), (256, 200, 4096), (3, 7, 50)])
def test_stein_diag(K, D, N):
    """(256, 200, ..) is the shape of BASELINE config C4-diagonal; the sums run as split-K matrix products (csrc/diag.cu)."""
    from gmmvi_b200 import ops
    g, X = make_problem(K, D, N, seed=7, diag=True)
    g32 = gmm32_of(g)
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), True)
    rng = np.random.default_rng(8)
    X64 = X.astype(np.float64)
    tgrad = rng.standard_normal((N, D)).astype(np.float32)
    mapping = np.sort(rng.integers(0, K, N)).astype(np.int32)
    mapping[-1] = K - 1
    lq64 = O.component_log_densities(g_in, X64)
    bg = O.logsumexp(lq64 + np.log(np.ones(K) / K)[:, None], axis=0).astype(np.float32)
    Href, gref = O.stein_ng(g_in, X64, mapping, bg.astype(np.float64), np.zeros(N), tgrad.astype(np.float64))
    lq = ops.logdens_diag(dev(X), dev(g32.means), dev(g32.chol_cov))
    logq = ops.mixture_lse(lq, dev(g32.log_weights))
    gq = ops.mixture_grad_diag(dev(X), dev(g32.means), dev(g32.chol_cov), lq, dev(g32.log_weights), logq)
    G = dev(tgrad) - gq
    iw = ops.importance_weights(lq, dev(bg), None, True, None, True)
    Hneg, gneg = ops.stein_diag(dev(X), dev(g32.means), dev(g32.chol_cov), iw["W"], G)
    assert rel_err(gneg.cpu().numpy(), gref) < NG_RTOL
    assert rel_err(Hneg.cpu().numpy(), Href) < NG_RTOL
    # the first-generation kernels (serial per (component, dimension)) agree
    import os
    os.environ["GMMVI_B200_DIAG_V1"] = "1"
    try:
        H1, g1 = ops.stein_diag(dev(X), dev(g32.means), dev(g32.chol_cov), iw["W"], G)
        lq1 = ops.logdens_diag(dev(X), dev(g32.means), dev(g32.chol_cov))
    finally:
        del os.environ["GMMVI_B200_DIAG_V1"]
    assert rel_err(Hneg.cpu().numpy(), H1.cpu().numpy()) < 2e-5 and rel_err(gneg.cpu().numpy(), g1.cpu().numpy()) < 2e-5
    assert rel_err(lq.cpu().numpy(), lq1.cpu().numpy()) < 1e-6


def _update_problem(K, D, seed, diag=False, last_eta=None):
    rng = np.random.default_rng(seed)
    g, _ = make_problem(K, D, 4, seed=seed, diag=diag)
    g32 = gmm32_of(g)
    if diag:
        H = rng.uniform(-0.2, 1.5, (K, D))
    else:
        A = rng.standard_normal((K, D, D))
        H = A @ A.transpose(0, 2, 1) / D - 0.3 * np.eye(D)       # indefinite but symmetric
    gn = rng.standard_normal((K, D)) * 0.5
    steps = rng.uniform(0.005, 0.5, K)
    g32.stepsizes = steps.astype(np.float32)
    if last_eta is not None:
        g32.last_log_etas = np.asarray(last_eta, np.float32)
    return g32, H.astype(np.float32), gn.astype(np.float32)


def _as64(g32):
    g = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                    g32.chol_cov.astype(np.float64), g32.diagonal_covs)
    g.stepsizes = g32.stepsizes.astype(np.float64)
    g.last_log_etas = g32.last_log_etas.astype(np.float64)
    g.num_received_updates = g32.num_received_updates.astype(np.float64)
    return g


@pytest.mark.parametrize("K,D", [(4, 3), (6, 10), (5, 20), (3, 100), (2, 256)])
@pytest.mark.parametrize("warm", [False, True])
def test_kl_update_full(K, D, warm):
    from gmmvi_b200 import ops
    last = np.full(K, 30.0) if warm else None
    g32, H, gn = _update_problem(K, D, seed=11 + D, last_eta=last)
    g64 = _as64(g32)
    traces = []
    info = O.kl_constrained_update(g64, H.astype(np.float64), gn.astype(np.float64), g64.stepsizes, 1.0, traces)
    om, oc, succ, etas, kls = ops.update_components("trust-region", False, dev(g32.means), dev(g32.chol_cov), dev(H),
                                                    dev(gn), dev(g32.stepsizes), dev(g32.last_log_etas), None, 1.0)
    assert np.array_equal(succ.cpu().numpy().astype(bool), info["success"])
    # eta is the outcome of a discrete bisection: it must match unless the oracle trace shows a near tie
    for k in range(K):
        near_tie = any(abs(abs(g64.stepsizes[k] - kl) - 0.1 * g64.stepsizes[k]) < 2e-3 * g64.stepsizes[k] or
                       abs(g64.stepsizes[k] - kl) < 1e-3 * g64.stepsizes[k] for _, kl in traces[k])
        if not near_tie:
            assert np.isclose(etas.cpu().numpy()[k], info["etas"][k], rtol=1e-5), (k, traces[k])
            assert rel_err(om.cpu().numpy()[k], g64.means[k]) < NG_RTOL
            cov = oc.cpu().numpy()[k].astype(np.float64)
            cov = cov @ cov.T
            ref = g64.chol_cov[k] @ g64.chol_cov[k].T
            assert rel_err(cov, ref) < NG_RTOL
            assert rel_err(oc.cpu().numpy()[k], g64.chol_cov[k]) < NG_RTOL
            if info["success"][k]:
                assert np.isclose(kls.cpu().numpy()[k], info["kls"][k], rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("D", [1, 2, 3, 5, 33, 100, 256])
@pytest.mark.parametrize("symmetric", [True, False])
def test_tridiag(D, symmetric):
    """Householder tridiagonalisation used by the KL-constrained update: T = P^T S P and h' = P^T h for the symmetric
    matrix S built from the UPPER triangle of B (the triangle the update kernel reads).  Orthogonal invariants: the
    spectrum, |h| and the Krylov moments h^T S^j h."""
    from gmmvi_b200 import ops
    rng = np.random.default_rng(50 + D)
    K = 3
    A = rng.standard_normal((K, D, D))
    B = A @ A.transpose(0, 2, 1) / D - 0.3 * np.eye(D) if symmetric else A
    h = rng.standard_normal((K, D))
    d, e, hp = (t.cpu().numpy().astype(np.float64) for t in ops.tridiag(dev(B.astype(np.float32)), dev(h.astype(np.float32))))
    for k in range(K):
        Bk = B[k].astype(np.float32).astype(np.float64)
        S = np.triu(Bk) + np.triu(Bk, 1).T
        T = np.diag(d[k]) + np.diag(e[k][:-1], 1) + np.diag(e[k][:-1], -1)
        scale = max(np.abs(np.linalg.eigvalsh(S)).max(), 1e-30)
        assert np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(S)).max() < 2e-5 * scale
        hk = h[k].astype(np.float32).astype(np.float64)
        assert abs(np.linalg.norm(hp[k]) - np.linalg.norm(hk)) < 1e-5 * np.linalg.norm(hk)
        for j in (1, 2, 3):
            ref = hk @ np.linalg.matrix_power(S, j) @ hk
            got = hp[k] @ np.linalg.matrix_power(T, j) @ hp[k]
            assert abs(got - ref) < 5e-5 * scale ** j * (hk @ hk)


@pytest.mark.parametrize("K,D,warm", [(6, 10, False), (4, 64, True), (3, 100, False), (3, 256, True), (2, 256, False)])
def test_kl_update_tridiagonal_search_matches_cholesky_search(K, D, warm, monkeypatch):
    """The bisection with KL(eta) from the tridiagonal form must take the decisions of the bisection that factors
    M(eta) for every eta: same eta (a discrete outcome), same evaluation count, same success flags, and new parameters
    equal to rounding (the final factorisation is the same code in both)."""
    from gmmvi_b200 import ops
    last = np.full(K, 30.0) if warm else None
    g32, H, gn = _update_problem(K, D, seed=91 + D, last_eta=last)
    args = ("trust-region", False, dev(g32.means), dev(g32.chol_cov), dev(H), dev(gn), dev(g32.stepsizes),
            dev(g32.last_log_etas), None, 1.0)
    monkeypatch.setenv("GMMVI_B200_UPDATE_TRIDIAG", "0")
    ref = [t.clone() for t in ops.update_components(*args)]
    ref_evals = ops.last_update_evals.clone()
    monkeypatch.setenv("GMMVI_B200_UPDATE_TRIDIAG", "1")
    new = ops.update_components(*args)
    assert torch.equal(ref[2], new[2])
    assert torch.equal(ref_evals, ops.last_update_evals)
    assert torch.equal(ref[3], new[3])                                  # etas
    assert torch.equal(ref[0], new[0]) and torch.equal(ref[1], new[1])  # same final factorisation -> same bits
    assert torch.allclose(ref[4], new[4], rtol=1e-6, atol=0)


@pytest.mark.parametrize("mode", ["direct", "iBLR"])
@pytest.mark.parametrize("K,D", [(5, 10), (3, 100), (2, 200)])
@pytest.mark.parametrize("first", [True, False])
def test_direct_iblr_update_full(mode, K, D, first):
    from gmmvi_b200 import ops
    g32, H, gn = _update_problem(K, D, seed=21 + D)
    g32.stepsizes = (g32.stepsizes * (0.05 if mode == "direct" else 1.0)).astype(np.float32)
    if not first:
        g32.num_received_updates = np.ones(K, np.float32)
    g64 = _as64(g32)
    fn = O.direct_update if mode == "direct" else O.iblr_update
    info = fn(g64, H.astype(np.float64), gn.astype(np.float64), g64.stepsizes)
    om, oc, succ, _, _ = ops.update_components(mode, False, dev(g32.means), dev(g32.chol_cov), dev(H), dev(gn),
                                               dev(g32.stepsizes), None, dev(g32.num_received_updates), 1.0)
    assert np.array_equal(succ.cpu().numpy().astype(bool), info["success"])
    assert rel_err(om.cpu().numpy(), g64.means) < NG_RTOL
    assert rel_err(oc.cpu().numpy(), g64.chol_cov) < NG_RTOL


def test_direct_update_rejects_non_pd():
    from gmmvi_b200 import ops
    K, D = 3, 8
    g32, H, gn = _update_problem(K, D, seed=31)
    H[1] = -50.0 * np.eye(D, dtype=np.float32)          # P + s R is not PD for component 1
    g32.stepsizes = np.full(K, 0.5, np.float32)
    g64 = _as64(g32)
    old_means, old_chol = g64.means.copy(), g64.chol_cov.copy()
    info = O.direct_update(g64, H.astype(np.float64), gn.astype(np.float64), g64.stepsizes)
    om, oc, succ, _, _ = ops.update_components("direct", False, dev(g32.means), dev(g32.chol_cov), dev(H), dev(gn),
                                               dev(g32.stepsizes), None, None, 1.0)
    s = succ.cpu().numpy().astype(bool)
    assert np.array_equal(s, info["success"]) and not s[1]
    assert np.array_equal(oc.cpu().numpy()[1], g32.chol_cov[1]) and np.array_equal(om.cpu().numpy()[1], g32.means[1])


@pytest.mark.parametrize("mode", ["trust-region", "iBLR"])
def test_update_diag(mode):
    from gmmvi_b200 import ops
    K, D = 7, 60
    g32, H, gn = _update_problem(K, D, seed=41, diag=True)
    g32.num_received_updates = np.array([0, 1, 0, 2, 1, 0, 3], np.float32)
    g64 = _as64(g32)
    if mode == "trust-region":
        info = O.kl_constrained_update(g64, H.astype(np.float64), gn.astype(np.float64), g64.stepsizes, 1.0)
    else:
        info = O.iblr_update(g64, H.astype(np.float64), gn.astype(np.float64), g64.stepsizes)
    om, oc, succ, etas, kls = ops.update_components(mode, True, dev(g32.means), dev(g32.chol_cov), dev(H), dev(gn),
                                                    dev(g32.stepsizes), dev(g32.last_log_etas),
                                                    dev(g32.num_received_updates), 1.0)
    assert np.array_equal(succ.cpu().numpy().astype(bool), info["success"])
    if mode == "trust-region":
        assert np.allclose(etas.cpu().numpy(), info["etas"], rtol=1e-5)
    assert rel_err(om.cpu().numpy(), g64.means) < NG_RTOL
    assert rel_err(oc.cpu().numpy(), g64.chol_cov) < NG_RTOL


@pytest.mark.parametrize("K", [1, 2, 50, 1500])
@pytest.mark.parametrize("tr", [True, False])
def test_weight_update(K, tr):
    from gmmvi_b200 import ops
    rng = np.random.default_rng(50 + K)
    w = rng.uniform(0.1, 1.0, K)
    logw = np.log(w / w.sum()).astype(np.float32)
    elr = (rng.standard_normal(K) * 3).astype(np.float32)
    g = O.OracleGMM(logw.copy(), np.zeros((K, 2), np.float32), np.ones((K, 2), np.float32), True)
    logw = g.log_weights.copy()
    trace = []
    if tr:
        O.trust_region_weight_update(g, elr, 0.05, 1.0, trace)
    else:
        O.direct_weight_update(g, elr, 0.3, 1.0)
    out, info = ops.weight_update(tr, dev(logw), dev(elr), 0.05 if tr else 0.3, 1.0)
    out = out.cpu().numpy().astype(np.float64)
    out = out - O.logsumexp(out)                      # GMM.replace_weights normalises (models/gmm.py:173-181)
    near_tie = any(abs(abs(0.05 - kl) - 0.005) < 1e-4 for _, kl in trace)
    if not near_tie:
        assert np.allclose(np.exp(out), np.exp(g.log_weights), rtol=2e-4, atol=1e-7)


def test_sampling_and_noise():
    from gmmvi_b200 import ops
    K, D = 5, 20
    g, _ = make_problem(K, D, 4, seed=60)
    g32 = gmm32_of(g)
    n_per = np.array([3, 0, 130, 1, 40], np.int32)
    N = int(n_per.sum())
    offsets = np.concatenate(([0], np.cumsum(n_per))).astype(np.int32)
    eps = ops.fill_normal(N, D, seed=1234, subsequence=7)
    e = eps.cpu().numpy()
    noise = lambda k, D_, n: e[offsets[k]:offsets[k + 1]].T
    Xref, mref = O.sample_from_components_no_shuffle(g32, n_per, noise)
    X, mapping = ops.sample_components(False, eps, dev(offsets, torch.int32), dev(g32.means), dev(g32.chol_cov),
                                       int(n_per.max()))
    assert np.array_equal(mapping.cpu().numpy(), mref)          # bit exact
    assert rel_err(X.cpu().numpy(), Xref) < 1e-6
    # the generator is counter based: a shard that starts at row r draws the same numbers
    part = ops.fill_normal(50, D, seed=1234, subsequence=7, row_offset=100)
    assert torch.equal(part, eps[100:150])
    big = ops.fill_normal(200000, 16, seed=5).cpu().numpy()
    assert abs(big.mean()) < 5e-3 and abs(big.std() - 1) < 5e-3
    # diagonal
    gd, _ = make_problem(K, D, 4, seed=61, diag=True)
    gd32 = gmm32_of(gd)
    Xref, _ = O.sample_from_components_no_shuffle(gd32, n_per, noise)
    X, mapping = ops.sample_components(True, eps, dev(offsets, torch.int32), dev(gd32.means), dev(gd32.chol_cov),
                                       int(n_per.max()))
    assert rel_err(X.cpu().numpy(), Xref) < 1e-6 and np.array_equal(mapping.cpu().numpy(), mref)


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_bgemm(ta, tb):
    from gmmvi_b200 import ops
    rng = np.random.default_rng(70)
    b, M, N, Kd = 3, 70, 45, 133
    A = rng.standard_normal((b, Kd, M) if ta else (b, M, Kd)).astype(np.float32)
    B = rng.standard_normal((b, N, Kd) if tb else (b, Kd, N)).astype(np.float32)
    ref = (A.transpose(0, 2, 1) if ta else A).astype(np.float64) @ (B.transpose(0, 2, 1) if tb else B).astype(np.float64)
    out = ops.bgemm(dev(A), dev(B), ta, tb, 0.5)
    assert rel_err(out.cpu().numpy(), 0.5 * ref) < 1e-5


@pytest.mark.parametrize("b,M,N,Kd,kseg,beta,lower", [
    (2, 300, 300, 1000, 4, 0.0, True),       # several segments, ragged tiles, lower triangle only
    (1, 517, 517, 128, 0, 1.0, True),        # trailing update of the blocked Cholesky: C -= T T^T
    (3, 130, 70, 260, 3, 1.0, False),        # rectangular, last segment shorter, K not a multiple of 32
    (1, 1030, 1030, 2052, 16, 0.0, True),    # more tiles than one wave of lower tiles per row
])
def test_tc_bgemm_segments_beta_lower(b, M, N, Kd, kseg, beta, lower):
    """gvi_tc_bgemm_ex_f32: round-to-nearest accumulated reduction segments, beta = 1, lower-triangle tile enumeration."""
    from gmmvi_b200 import ops
    rng = np.random.default_rng(M + Kd)
    same = M == N
    A = rng.standard_normal((b, M, Kd)).astype(np.float32)
    B = A if same else rng.standard_normal((b, N, Kd)).astype(np.float32)
    C0 = rng.standard_normal((b, M, N)).astype(np.float32)
    alpha = -1.0 if beta else 0.5
    ref = alpha * (A.astype(np.float64) @ B.astype(np.float64).transpose(0, 2, 1)) + beta * C0
    C = dev(C0.copy())
    ops.bgemm_ex(dev(A), dev(B), C, False, True, alpha, beta, kseg, lower)
    out = C.cpu().numpy()
    rows, cols = np.arange(M)[:, None], np.arange(N)[None, :]
    written = np.ones((M, N), bool) if not lower else (cols // 256) * 256 <= (rows // 128) * 128 + 127
    assert written[np.tril_indices(M, 0, N)].all()
    scale = np.abs(ref).max()
    assert np.abs(out - ref)[:, written].max() < 2e-6 * scale * max(1.0, np.sqrt(Kd / 256))
    assert np.array_equal(out[:, ~written], C0[:, ~written])        # skipped tiles are not touched
    # A^T B through the transposing split as well (the form the estimator's normal equations have)
    if not beta:
        C2 = dev(C0.copy())
        At = np.ascontiguousarray(A.transpose(0, 2, 1))
        ops.bgemm_ex(dev(At), dev(np.ascontiguousarray(B.transpose(0, 2, 1))), C2, True, False, alpha, 0.0, kseg, lower)
        assert np.array_equal(C2.cpu().numpy()[:, written], out[:, written])


@pytest.mark.parametrize("b,M,N,Kd,kseg,beta,lower", [
    (2, 300, 300, 1000, 4, 0.0, True),
    (1, 517, 517, 128, 0, 1.0, True),
    (3, 130, 70, 264, 3, 1.0, False),
    (1, 1030, 1030, 2056, 8, 0.0, True),
])
def test_tc_bgemm_h16_split(b, M, N, Kd, kseg, beta, lower):
    """gvi_tc_bgemm_h16_f32: the batched product in the 2 x fp16 split precision; rows of very different magnitude
    (a factor 2^-12 .. 1 per row) check that the per-matrix power-of-two scale keeps fp32-grade accuracy."""
    from gmmvi_b200 import ops
    rng = np.random.default_rng(M + Kd + 1)
    same = M == N
    rowscale = 2.0 ** rng.integers(-12, 1, size=(b, M, 1))
    A = (rng.standard_normal((b, M, Kd)) * rowscale).astype(np.float32)
    B = A if same else rng.standard_normal((b, N, Kd)).astype(np.float32)
    C0 = rng.standard_normal((b, M, N)).astype(np.float32)
    alpha = -1.0 if beta else 0.5
    prod = A.astype(np.float64) @ B.astype(np.float64).transpose(0, 2, 1)
    ref = alpha * prod + beta * C0
    # error scale of an fp32-grade product: |a|.|b| per entry
    mag = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).transpose(0, 2, 1) + beta * np.abs(C0)
    C = dev(C0.copy())
    Ad = dev(A)
    ops.bgemm_h16(Ad, Ad if same else dev(B), C, alpha, beta, kseg, lower)
    out = C.cpu().numpy()
    rows, cols = np.arange(M)[:, None], np.arange(N)[None, :]
    written = np.ones((M, N), bool) if not lower else (cols // 256) * 256 <= (rows // 128) * 128 + 127
    err = np.abs(out - ref) / (mag + 1e-30)
    assert err[:, written].max() < 8e-6, err[:, written].max()       # incl. the truncating accumulator (<= 96 adds per segment)
    assert np.array_equal(out[:, ~written], C0[:, ~written])


@pytest.fixture(params=["h16", "h16-single-panels", "tf32", "simt"])
def more_route(request, monkeypatch):
    monkeypatch.setenv("GMMVI_B200_MORE_TC", {"h16": "1", "h16-single-panels": "1", "tf32": "tf32", "simt": "0"}[request.param])
    monkeypatch.setenv("GMMVI_B200_MORE_PAIRS", "0" if request.param == "h16-single-panels" else "1")
    # the column-by-column diagonal-block kernel rides along with the single-panel route, the blocked one with the rest
    monkeypatch.setenv("GMMVI_B200_MORE_POTRF", "columns" if request.param == "h16-single-panels" else "blocked")
    from gmmvi_b200 import _lib
    assert _lib.lib().gvi_more_tensor_cores() == {"h16": 2, "h16-single-panels": 2, "tf32": 1, "simt": 0}[request.param]
    return request.param


@pytest.mark.parametrize("K,D,N,self_norm", [(3, 4, 300, True), (4, 12, 1500, True), (2, 20, 2000, False),
                                              (2, 40, 4000, True), (3, 30, 2501, True)])
def test_more_estimator(K, D, N, self_norm, more_route):
    """MORE against the oracle (ng_estimator.py:296-376): needs N >= F = D(D+1)/2 + D + 1 samples.  Both routes: the
    normal equations / Cholesky trailing updates on the tensor cores (3xTF32, pre-split transposed features) and in the
    SIMT fp32 engine; N = 2501 is not a multiple of 4 (zero-padded reduction), D = 30 gives F + 1 = 497 (pitch 500)."""
    from gmmvi_b200 import ops
    g, X = make_problem(K, D, N, seed=80 + D, scale=1.0)
    g32 = gmm32_of(g)
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), False, initial_regularizer=1e-8)
    rng = np.random.default_rng(81)
    X64 = X.astype(np.float64)
    # a smooth target: quadratic + small non-quadratic perturbation
    Q = rng.standard_normal((D, D)); Q = Q @ Q.T / D + np.eye(D)
    tl = (-0.5 * np.einsum("ni,ij,nj->n", X64, Q, X64) + 0.1 * np.sin(X64).sum(1)).astype(np.float32)
    mapping = np.sort(rng.integers(0, K, N)).astype(np.int32)
    mapping[-1] = K - 1
    lq64 = O.component_log_densities(g_in, X64)
    bg = O.logsumexp(lq64 + g_in.log_weights[:, None], axis=0).astype(np.float32)
    Href, gref = O.more_ng(g_in, X64, mapping, bg.astype(np.float64), tl.astype(np.float64), None, False, self_norm)
    linv, prec, cst, _ = ops.prepare_full(dev(g32.chol_cov))
    lq = ops.logdens_full(dev(X), dev(g32.means), linv, cst)
    logq = ops.mixture_lse(lq, dev(g32.log_weights))
    iw = ops.importance_weights(lq, dev(bg), None, self_norm, None, True)
    y = dev(tl) - logq
    l2 = torch.full((K,), 1e-8, device="cuda")
    quad, lin, ok = ops.more_fit(l2, dev(X), y, iw["W"], dev(g32.means), linv, memory_budget_bytes=64 << 20)
    assert ok.cpu().numpy().all()
    gneg = (quad @ dev(g32.means).unsqueeze(2)).squeeze(2) - lin
    assert rel_err(quad.cpu().numpy(), Href) < 5e-4, rel_err(quad.cpu().numpy(), Href)
    assert rel_err(gneg.cpu().numpy(), gref) < 5e-4


@pytest.mark.parametrize("n1,n2,D", [(1, 1, 1), (70, 33, 5), (300, 257, 20), (1100, 500, 40)])
def test_mmd_matches_oracle(n1, n2, D):
    """experiments/evaluation/mmd.py on device: bandwidth (an element of the data: exact), U-statistics and MMD against
    the restatement; n1 = 1100 exercises the 1000-point cap of the median trick and several CTA tiles."""
    from gmmvi_b200.experiments.evaluation.mmd import MMD
    rng = np.random.default_rng(n1 + D)
    G = (rng.standard_normal((n1, D)) * 2 + 1).astype(np.float32)
    S = (rng.standard_normal((n2, D)) * 2.2 + 1.3).astype(np.float32)
    if n1 == 1:
        G[:] = 0.5                                   # a single point: sigma = 0 -> inf bandwidth like the reference
        m = MMD(G, 3.0)
        assert float(torch.diagonal(m.sigma)[0]) == 0.0
        return
    alpha = 20.0
    m = MMD(G, alpha)
    sig = O.mmd_sigma(G, dtype=np.float32)
    assert np.array_equal(torch.diagonal(m.sigma).cpu().numpy(), sig)
    assert m.sigma.shape == (D, D)
    ref_u = O.mmd_kernel_sum(S, S, sig, alpha, np.float64)
    ref_mix = O.mmd_kernel_sum(G, S, sig, alpha, np.float64)
    assert np.isclose(float(m.compute_ustat(dev(S), alpha)), ref_u, rtol=2e-5)
    assert np.isclose(float(m.kernel_mix(dev(S), alpha)), ref_mix, rtol=2e-5, atol=1e-12)
    ref = O.mmd(G, S, alpha, np.float64)
    assert np.isclose(float(m.compute_MMD(dev(S))), ref, rtol=1e-3, atol=1e-6)
    assert abs(float(m.compute_MMD(dev(G)))) < 1e-6


@pytest.mark.parametrize("K,D,N", [(1, 1, 7), (37, 8, 1000), (19, 9, 515), (45, 20, 9000), (128, 10, 3001), (11, 16, 129),
                                   (9, 17, 400), (10, 24, 640), (13, 25, 300), (8, 31, 777), (12, 32, 1500)])
def test_small_dimension_kernels(K, D, N, monkeypatch):
    """The D <= 32 kernels (csrc/small_dim.cu: the shapes of BASELINE configs C1 / C2 -- D = 20 and D = 10) against the
    fp64 oracle and against the general tile engines (GMMVI_B200_SMALL_DIM=0) on the same inputs: component log
    densities, mixture gradient, Stein statistics with weightless blocks skipped."""
    from gmmvi_b200 import ops

    def run():
        ops.clear_caches()
        g, X = make_problem(K, D, N, seed=21, scale=4.0)
        g32 = gmm32_of(g)
        rng = np.random.default_rng(22)
        tgrad = rng.standard_normal((N, D)).astype(np.float32)
        linv, prec, cst, _ = ops.prepare_full(dev(g32.chol_cov))
        lq = ops.logdens_full(dev(X), dev(g32.means), linv, cst, memo=False, tensor_cores=False)
        logq = ops.mixture_lse(lq, dev(g32.log_weights))
        gq = ops.mixture_grad_full(dev(X), dev(g32.means), prec, lq, dev(g32.log_weights), logq, tensor_cores=False)
        bg = ops.mixture_lse(lq, dev(np.log(np.ones(K, np.float32) / K)))
        iw = ops.importance_weights(lq, bg, None, True, None, True, False, False, True)
        G = (dev(tgrad) - gq).contiguous()
        H, gn = ops.stein_full(dev(X), dev(g32.means), prec, iw["W"], iw["active"], G, True)
        return g32, X, tgrad, [t.cpu().numpy() for t in (lq, gq, H, gn, bg)]
    g32, X, tgrad, small = run()
    monkeypatch.setenv("GMMVI_B200_SMALL_DIM", "0")
    _, _, _, general = run()
    monkeypatch.delenv("GMMVI_B200_SMALL_DIM")
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64), g32.chol_cov.astype(np.float64), False)
    X64 = X.astype(np.float64)
    ref_lq = O.component_log_densities(g_in, X64)
    _, ref_grad, _ = O.log_density_and_grad(g_in, X64)
    mapping = np.zeros(N, np.int32); mapping[-1] = K - 1
    Href, gref = O.stein_ng(g_in, X64, mapping, small[4].astype(np.float64), np.zeros(N), tgrad.astype(np.float64), False, True)
    assert rel_err(small[0], ref_lq) < LOGDENS_RTOL
    assert np.max(np.abs(small[0] - ref_lq) / np.maximum(np.abs(ref_lq), 1.0)) < LOGDENS_RTOL        # element-wise
    assert rel_err(small[1], ref_grad) < NG_RTOL
    assert rel_err(small[2], Href) < NG_RTOL and rel_err(small[3], gref) < NG_RTOL
    for a, b in zip(small[:4], general[:4]):
        assert rel_err(a, b) < 2e-5


def test_construction_time_cholesky():
    """gvi_cholesky_f32 (models/full_cov_gmm.py:23, :67): fp64 arithmetic on the device, NaN factor + ok = 0 for a matrix
    that is not positive definite (tf.linalg.cholesky's behaviour)."""
    from gmmvi_b200 import ops
    rng = np.random.default_rng(3)
    for K, D in ((5, 3), (7, 20), (3, 256), (1, 1)):
        A = rng.standard_normal((K, D, D))
        cov = (A @ A.transpose(0, 2, 1) / D + np.eye(D)).astype(np.float32)
        L, ok = ops.cholesky(dev(cov))
        ref = np.linalg.cholesky(cov.astype(np.float64))
        assert ok.cpu().numpy().all()
        assert rel_err(L.cpu().numpy(), ref) < 2e-7
        assert np.all(np.triu(L.cpu().numpy(), 1) == 0)
    bad = np.stack([np.eye(4, dtype=np.float32), -np.eye(4, dtype=np.float32)])
    L, ok = ops.cholesky(dev(bad))
    assert ok.cpu().numpy().tolist() == [1, 0] and np.isnan(L[1].cpu().numpy()).all() and np.isfinite(L[0].cpu().numpy()).all()


@pytest.mark.parametrize("K,D,mode", [(5, 12, "direct"), (4, 64, "direct"), (3, 130, "direct"), (2, 256, "direct"),
                                      (5, 12, "iBLR"), (3, 96, "iBLR")])
def test_update_with_nonsymmetric_hessian_matches_oracle(K, D, mode):
    """gvi_update_full_general_f32: the direct / iBLR update for a NON-symmetric -E[H] (what Stein with standard
    importance weights hands over, ng_estimator.py:168): general inverse + Cholesky of its lower triangle, restated
    literally (ng_based_component_updater.py:116-118, 199-200); one component is made to fail (not positive definite)."""
    from gmmvi_b200 import ops
    g32, H, gn = _update_problem(K, D, seed=40 + D)
    rng = np.random.default_rng(41 + D)
    H = H + 0.15 * rng.standard_normal(H.shape).astype(np.float32) / np.sqrt(D)        # asymmetric part
    steps = np.full(K, 0.05 if mode == "direct" else 0.02, np.float32)
    if mode == "direct":          # (the iBLR precision P + s R + s^2/2 R Sigma R is positive definite for any symmetric R)
        H[-1] = -40.0 * np.eye(D, dtype=np.float32)                                     # P + s R not positive definite
        steps[-1] = 1.0
    g32.stepsizes = steps
    g32.num_received_updates = np.array([0.0] + [1.0] * (K - 1), np.float32)           # iBLR: first update keeps the mean
    g64 = _as64(g32)
    info = (O.direct_update if mode == "direct" else O.iblr_update)(g64, H.astype(np.float64), gn.astype(np.float64),
                                                                    steps.astype(np.float64))
    _, prec, _, _ = ops.prepare_full(dev(g32.chol_cov))
    om, oc, succ = ops.update_components_general(mode, dev(g32.means), dev(g32.chol_cov), prec, dev(H), dev(gn), dev(steps),
                                                 dev(g32.num_received_updates))
    assert np.array_equal(succ.cpu().numpy().astype(bool), info["success"])
    assert rel_err(om.cpu().numpy(), g64.means) < NG_RTOL
    assert rel_err(oc.cpu().numpy(), g64.chol_cov) < NG_RTOL
    if mode == "direct":          # the rejected component keeps its parameters bit for bit
        assert not info["success"][-1]
        assert np.array_equal(om.cpu().numpy()[-1], g32.means[-1]) and np.array_equal(oc.cpu().numpy()[-1], g32.chol_cov[-1])
