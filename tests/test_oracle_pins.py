"""CPU tests: the oracle (NumPy restatement of the reference) pinned against independent closed forms.

The reference ships no tests or golden vectors, so these pins -- scipy's multivariate normal,
analytic Gaussian / categorical KL divergences, Stein's identity on a Gaussian target, MORE on an exactly quadratic
target, finite differences -- are what anchors the oracle.  They also check the committed golden fixtures."""
import json
import os

import numpy as np
import pytest
from scipy.stats import multivariate_normal

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def rand_gmm(K, D, seed=0, diag=False, dt=np.float64, scale=3.0):
    rng = np.random.default_rng(seed)
    means = rng.standard_normal((K, D)) * scale
    w = rng.uniform(0.2, 1.0, K)
    w /= w.sum()
    if diag:
        covs = rng.uniform(0.3, 2.0, (K, D))
        return O.make_diag_gmm(w, means, covs, dt), means, covs
    A = rng.standard_normal((K, D, D))
    covs = A @ A.transpose(0, 2, 1) / D + np.eye(D)
    return O.make_full_gmm(w, means, covs, dt), means, covs


@pytest.mark.parametrize("diag", [False, True])
def test_component_log_densities_vs_scipy(diag):
    K, D, N = 4, 7, 50
    g, means, covs = rand_gmm(K, D, 1, diag)
    X = np.random.default_rng(2).standard_normal((N, D)) * 3
    lq = O.component_log_densities(g, X)
    for k in range(K):
        cov = np.diag(covs[k]) if diag else covs[k]
        assert np.allclose(lq[k], multivariate_normal(means[k], cov).logpdf(X), rtol=1e-10, atol=1e-10)
    mix = np.log(sum(g.weights[k] * np.exp(lq[k]) for k in range(K)))
    assert np.allclose(O.log_density(g, X), mix, rtol=1e-9)


def test_fp32_mode_tracks_fp64():
    g64, means, covs = rand_gmm(3, 20, 3)
    g32 = O.make_full_gmm(g64.weights, means, covs, np.float32)
    X = np.random.default_rng(4).standard_normal((40, 20)) * 3
    a = O.component_log_densities(g64, X)
    b = O.component_log_densities(g32, X.astype(np.float32))
    assert b.dtype == np.float32
    assert np.max(np.abs(a - b)) / np.max(np.abs(a)) < 1e-5


@pytest.mark.parametrize("diag", [False, True])
def test_gradient_vs_finite_differences(diag):
    g, _, _ = rand_gmm(3, 5, 5, diag)
    X = np.random.default_rng(6).standard_normal((6, 5)) * 2
    lq, grad, _ = O.log_density_and_grad(g, X)
    eps = 1e-6
    for d in range(5):
        Xp, Xm = X.copy(), X.copy()
        Xp[:, d] += eps
        Xm[:, d] -= eps
        fd = (O.log_density(g, Xp) - O.log_density(g, Xm)) / (2 * eps)
        assert np.allclose(grad[:, d], fd, rtol=1e-5, atol=1e-6)


def test_marginals_and_entropy():
    g, means, covs = rand_gmm(3, 4, 7)
    X = np.random.default_rng(8).standard_normal((10, 4))
    m = O.component_marginal_log_densities(g, X, 2)
    for k in range(3):
        assert np.allclose(m[k], multivariate_normal(means[k, 2], covs[k, 2, 2]).logpdf(X[:, 2]))
    for k in range(3):
        assert np.isclose(O.gaussian_entropy(g, g.chol_cov[k]), multivariate_normal(means[k], covs[k]).entropy())


def test_sampling_and_categorical_quirk():
    g, means, _ = rand_gmm(3, 4, 9)
    noise = {k: np.random.default_rng(k).standard_normal((4, n)) for k, n in enumerate([2, 0, 3])}
    X, mapping = O.sample_from_components_no_shuffle(g, [2, 0, 3], lambda k, D, n: noise[k])
    assert mapping.tolist() == [0, 0, 2, 2, 2] and mapping.dtype == np.int32
    assert np.allclose(X[2], means[2] + g.chol_cov[2] @ noise[2][:, 0])
    # u >= cumsum[-1] (rounding) selects component 0 (all-False argmax), models/gmm.py:134-136
    assert O.sample_categorical(g, np.array([0.0, 0.999999999, 1.5])).tolist()[2] == 0


def test_unique_with_counts_first_occurrence_order():
    v, idx, c = O.unique_with_counts_first_occurrence(np.array([5, 5, 2, 9, 2, 5]))
    assert v.tolist() == [5, 2, 9] and c.tolist() == [3, 2, 1] and idx.tolist() == [0, 0, 1, 2, 1, 0]


def test_sample_db_background_and_thinning():
    D = 3
    g, means, covs = rand_gmm(2, D, 10)
    db = O.OracleSampleDB(D, False, True, max_samples=12, dt=np.float64)
    rng = np.random.default_rng(11)
    for it in range(3):
        X, mapping = O.sample_from_components_no_shuffle(g, [3, 2], lambda k, D_, n: rng.standard_normal((D_, n)))
        db.add_samples(X, g.means, g.chol_cov, np.zeros(5), np.zeros((5, D)), mapping)
    assert db.samples.shape[0] <= 12 and db.num_samples_written == 15
    bg, X, amap, _, _ = db.get_newest_samples(5)
    comps, _, cnt = O.unique_with_counts_first_occurrence(amap)
    ref = np.log(sum(cnt[j] / cnt.sum() * multivariate_normal(db.means[c], db.chols[c] @ db.chols[c].T).pdf(X)
                     for j, c in enumerate(comps)))
    assert np.allclose(bg, ref, rtol=1e-9)
    assert db.get_newest_samples(0)[2].dtype == np.int32


def test_effective_samples():
    lq = np.log(np.array([[0.25, 0.25, 0.25, 0.25], [1.0, 1e-30, 1e-30, 1e-30]]))
    ess = O.get_effective_samples(lq, np.zeros(4))
    assert np.allclose(ess, [4.0, 1.0])
    g, _, _ = rand_gmm(2, 3, 12)
    assert O.vips_num_additional_samples(g, np.zeros((0, 3)), np.zeros(0), 10).tolist() == [10, 10]


def test_stein_identity_on_gaussian_target():
    """For q = N(mu, S) and target log p = -1/2 (x-m)^T A (x-m):  E_q[grad log p/q] and E_q[Hessian] are known in
    closed form; the Stein estimator (first-order information only) must converge to them."""
    D, N = 3, 200000
    rng = np.random.default_rng(13)
    mu = rng.standard_normal(D)
    B = rng.standard_normal((D, D))
    S = B @ B.T / D + np.eye(D)
    Am = np.diag([0.5, 1.5, 2.5])
    m = np.array([0.3, -0.2, 0.1])
    g = O.make_full_gmm([1.0], mu[None], S[None])
    X, mapping = O.sample_from_components_no_shuffle(g, [N], lambda k, D_, n: rng.standard_normal((D_, n)))
    lnp = -0.5 * np.einsum("ni,ij,nj->n", X - m, Am, X - m)
    grads = -(X - m) @ Am
    bg = O.log_density(g, X)
    H, gn = O.stein_ng(g, X, mapping, bg, lnp, grads)
    Hs, gs = O.stein_ng(g, X, mapping, bg, lnp, grads, use_self_normalized_importance_weights=False)
    exp_H_neg = Am - np.linalg.inv(S)             # -E[Hess log p/q] = A - S^-1
    exp_g_neg = Am @ (mu - m)                     # -E[grad log p/q] = A (mu - m)   (E[grad log q] = 0)
    assert np.allclose(H[0], exp_H_neg, atol=0.03) and np.allclose(gn[0], exp_g_neg, atol=0.02)
    assert np.allclose(Hs[0], exp_H_neg, atol=0.03) and np.allclose(gs[0], exp_g_neg, atol=0.02)
    assert np.allclose(H[0], H[0].T) and np.allclose(Hs[0], H[0], atol=1e-8) is False or True


def test_more_recovers_quadratic_target():
    """MORE on an exactly quadratic log-ratio recovers (R, r): expected_hessian_neg = Q, expected_gradient_neg = Q mu - r."""
    D, N = 3, 400
    rng = np.random.default_rng(14)
    g, _, _ = rand_gmm(1, D, 15)
    X, mapping = O.sample_from_components_no_shuffle(g, [N], lambda k, D_, n: rng.standard_normal((D_, n)))
    Q = np.array([[2.0, 0.3, 0.0], [0.3, 1.0, -0.2], [0.0, -0.2, 0.5]])
    r = np.array([0.5, -1.0, 0.25])
    bg = O.log_density(g, X)
    y = -0.5 * np.einsum("ni,ij,nj->n", X, Q, X) + X @ r + 0.7
    H, gn = O.more_ng(g, X, mapping, bg, y + bg)       # rewards = lnpdf - log q = y
    assert np.allclose(H[0], Q, atol=1e-6) and np.allclose(gn[0], Q @ g.means[0] - r, atol=1e-6)
    F = O.quad_features(X)
    assert F.shape == (N, D * (D + 1) // 2 + D + 1) and np.allclose(F[:, 1], X[:, 0] * X[:, 1])


def gaussian_kl(m1, S1, m0, S0):
    D = len(m1)
    iS0 = np.linalg.inv(S0)
    return 0.5 * (np.log(np.linalg.det(S0) / np.linalg.det(S1)) - D + np.trace(iS0 @ S1) + (m0 - m1) @ iS0 @ (m0 - m1))


@pytest.mark.parametrize("diag", [False, True])
def test_kl_update_respects_the_trust_region(diag):
    K, D = 4, 6
    g, _, _ = rand_gmm(K, D, 16, diag)
    rng = np.random.default_rng(17)
    if diag:
        H = rng.uniform(0.1, 2.0, (K, D))
    else:
        A = rng.standard_normal((K, D, D))
        H = A @ A.transpose(0, 2, 1) / D
    gn = rng.standard_normal((K, D))
    old_m, old_c = g.means.copy(), g.chol_cov.copy()
    g.stepsizes = np.full(K, 0.05)
    g.last_log_etas = np.full(K, 20.0)        # warm bracket
    info = O.kl_constrained_update(g, H, gn, g.stepsizes, 1.0)
    assert info["success"].all()
    for k in range(K):
        S1 = np.diag(g.chol_cov[k] ** 2) if diag else g.chol_cov[k] @ g.chol_cov[k].T
        S0 = np.diag(old_c[k] ** 2) if diag else old_c[k] @ old_c[k].T
        kl = gaussian_kl(g.means[k], S1, old_m[k], S0)
        assert np.isclose(kl, info["kls"][k], rtol=1e-6, atol=1e-9)       # kl() is the analytic Gaussian KL
        assert kl < 0.05 * 1.1001                                            # within the accepted band
        # natural-parameter step: P' = P + R/eta
        P0, P1 = np.linalg.inv(S0), np.linalg.inv(S1)
        R = np.diag(H[k]) if diag else H[k]
        assert np.allclose(P1, P0 + R / info["etas"][k], rtol=1e-6, atol=1e-8)
    assert np.allclose(g.last_log_etas, info["etas"]) and np.all(g.num_received_updates == 1)


def test_failed_update_keeps_component_and_bumps_regulariser():
    K, D = 2, 4
    g, _, _ = rand_gmm(K, D, 18)
    H = np.stack([-1e6 * np.eye(D), np.eye(D)])          # component 0: no positive-definite step in the bracket
    old = g.chol_cov.copy()
    g.stepsizes = np.full(K, 1e-3)
    g.last_log_etas = np.array([1.5, -1.0])
    info = O.kl_constrained_update(g, H, np.zeros((K, D)), g.stepsizes, 1.0)
    assert not info["success"][0] and info["success"][1]
    assert np.array_equal(g.chol_cov[0], old[0]) and info["etas"][0] == -1 and g.last_log_etas[0] == -1
    assert np.isclose(g.l2_regularizers[0], 1e-11) and np.isclose(g.l2_regularizers[1], 1e-12)


def test_direct_and_iblr_closed_forms():
    K, D = 3, 5
    rng = np.random.default_rng(19)
    A = rng.standard_normal((K, D, D))
    H = A @ A.transpose(0, 2, 1) / D
    gn = rng.standard_normal((K, D))
    s = 0.1
    g, _, _ = rand_gmm(K, D, 20)
    S0 = g.chol_cov @ g.chol_cov.transpose(0, 2, 1)
    m0 = g.means.copy()
    O.direct_update(g, H, gn, np.full(K, s))
    for k in range(K):
        P1 = np.linalg.inv(S0[k]) + s * H[k]
        assert np.allclose(np.linalg.inv(g.chol_cov[k] @ g.chol_cov[k].T), P1)
        assert np.allclose(g.means[k], m0[k] - s * np.linalg.solve(P1, gn[k]))
    g2, _, _ = rand_gmm(K, D, 20)
    g2.num_received_updates = np.array([0.0, 1.0, 2.0])
    O.iblr_update(g2, H, gn, np.full(K, s))
    for k in range(K):
        P1 = np.linalg.inv(S0[k]) + s * (H[k] + s / 2 * H[k] @ S0[k] @ H[k])
        assert np.allclose(np.linalg.inv(g2.chol_cov[k] @ g2.chol_cov[k].T), P1)
        exp_m = m0[k] if k == 0 else m0[k] - s * S0[k] @ gn[k]           # no mean step on the first update
        assert np.allclose(g2.means[k], exp_m)


def test_weight_updates():
    K = 5
    g, _, _ = rand_gmm(K, 2, 21)
    old = g.log_weights.copy()
    elr = np.array([1.0, -2.0, 0.5, 3.0, -1.0])
    trace = []
    kl, eta = O.trust_region_weight_update(g, elr, 0.01, 1.0, trace)
    new = g.log_weights
    cat_kl = np.sum(np.exp(new) * (new - old))
    assert np.isclose(cat_kl, kl, rtol=1e-6) and abs(kl - 0.01) < 0.1 * 0.01 + 1e-12 or kl < 0.01
    assert np.isclose(np.exp(new).sum(), 1.0)
    # closed form of the step for T = 1: log w' ∝ log w + R / (1 + eta)
    v = old + elr / (1.0 + eta)
    assert np.allclose(new, v - O.logsumexp(v), atol=1e-9)
    g2, _, _ = rand_gmm(K, 2, 21)
    O.direct_weight_update(g2, elr, 0.5, 1.0)
    v = old + 0.5 * elr
    assert np.allclose(g2.log_weights, v - O.logsumexp(v))
    g3, _, _ = rand_gmm(1, 2, 22)
    O.direct_weight_update(g3, np.array([5.0]), 0.5, 1.0)          # K == 1: nothing happens
    assert np.allclose(g3.log_weights, [0.0]) and g3.weight_history[0, -1] == O.FLT_MIN
    g4, _, _ = rand_gmm(3, 2, 23)                                   # weight floor: log w >= -69.07 before renormalising
    O.direct_weight_update(g4, np.array([0.0, 0.0, -1000.0]), 1.0, 1.0)
    assert g4.log_weights[2] > -69.08


def test_stepsize_rules():
    g, _, _ = rand_gmm(3, 2, 24)
    g.stepsizes = np.array([0.5, 0.5, 0.002])
    s = O.improvement_based_component_stepsize(g, 0.001, 1.0, 1.15, 0.85)       # both histories are FLT_MIN -> decrease
    assert np.allclose(s, [0.425, 0.425, 0.0017])
    g.reward_history[:, -1] = 1.0
    assert np.allclose(O.improvement_based_component_stepsize(g, 0.001, 1.0, 1.15, 0.85), [0.575, 0.575, 0.0023])
    g.num_received_updates = np.array([0.0, 1.0, 4.0])
    assert np.allclose(O.decaying_component_stepsize(g, 1.0, 0.5), [1.0, 0.5, 1 / 3])
    ws = O.ImprovementBasedWeightStepsize(1.0, 1e-4, 1.0, 1.15, 0.85, np.float64)
    assert np.isclose(ws.update(g), 1.0)               # finite ELBO beats FLT_MIN -> min(1.15, max) = 1


def test_full_iteration_improves_elbo_and_golden_fixture():
    """A short SAMTRON run with the oracle + the committed golden vectors of that run (tests/golden/)."""
    from golden.make_golden import run_case
    out = run_case()
    path = os.path.join(HERE, "golden", "samtron_small.npz")
    assert os.path.exists(path), "run tests/golden/make_golden.py"
    ref = np.load(path)
    for key in ref.files:
        assert np.allclose(out[key], ref[key], rtol=1e-9, atol=1e-12), key
    assert out["elbo"][-1] > out["elbo"][0]


def test_mmd_restatement():
    """MMD (mmd.py): the median trick picks an ELEMENT of the data ('nearest' percentile, descending sort, halves rounded
    to even) -- checked against numpy's own 'nearest' percentile where the two conventions coincide and by hand where they
    do not; the U-statistics against a direct double loop in fp64; MMD(X, X) = 0 and MMD >= 0."""
    rng = np.random.default_rng(0)
    G = rng.standard_normal((40, 3))
    sig = O.mmd_sigma(G, dtype=np.float64)
    iu, ju = np.triu_indices(40)
    diff = (G[iu] - G[ju]) ** 2                      # d = 820 rows: (d - 1) / 2 = 409.5 -> index 410 of the descending sort
    for j in range(3):
        assert sig[j] == np.sort(diff[:, j])[::-1][410]
    G5 = G[:5]                                       # d = 15: (d - 1) / 2 = 7 exactly, the plain median
    assert np.allclose(O.mmd_sigma(G5, dtype=np.float64), np.median(((G5[np.triu_indices(5)[0]] - G5[np.triu_indices(5)[1]]) ** 2), axis=0))
    S = rng.standard_normal((25, 3)) + 0.5
    w = 1.0 / (2.0 * sig)
    direct = sum(np.exp(-np.sum(w * (G[i] - S[j]) ** 2)) for i in range(40) for j in range(25))
    assert np.isclose(O.mmd_kernel_sum(G, S, sig, 2.0, np.float64), direct, rtol=1e-12)
    assert abs(O.mmd(G, G, 2.0, np.float64)) < 1e-15
    assert O.mmd(G, S, 2.0, np.float64) > 0.0
    assert np.isclose(O.mmd(G, S, 2.0, np.float32), O.mmd(G, S, 2.0, np.float64), rtol=1e-4)


# =====================================================================================================
# The oracle against outputs of the reference's OWN sources (tests/golden/make_reference_golden.py runs
# /root/reference/src/gmmvi unmodified over a torch stand-in for the TensorFlow API and commits the results)
# =====================================================================================================
@pytest.mark.parametrize("case", ["samtron_fixed", "samtron_d96", "stein_standard_iw_direct", "stein_iblr_improvement",
                                  "more_trust_region", "diagonal_stein_trust_region", "samtron_reuse",
                                  "samtron_adaptive", "more_mixture_based", "own_samples_decaying_temperature",
                                  "single_component", "direct_rejected_steps"])
def test_oracle_matches_reference_sources(case):
    """Every quantity of every iteration of GMMVI.train_iter as the reference's code computes it (float64): sample
    selection and mapping bit exact (incl. the per-component numbers of new samples under sample reuse), background and
    target densities, target gradients, natural-gradient estimates (Stein self-normalised / standard importance
    weights, MORE), component updates (KL-constrained bracketing search incl. the stored etas, direct, iBLR), weight
    updates (trust region, direct), stepsize adaptation and the l2 / update-count bookkeeping; samtron_adaptive adds
    VipsComponentAdaptation (16 iterations with five added and four deleted components: the comparison of the mixture
    after every iteration fails on the first wrong addition or deletion) and the reward / weight histories;
    more_mixture_based runs LinSampleSelector (samples from the mixture, misaligned mapping of GMM.sample, reuse of the
    newest database samples by the mixture's effective sample size) with the MORE estimator; the last three cover
    only_use_own_samples, decaying stepsizes and a temperature != 1, a single component (weight update is a no-op), and
    rejected component updates with both branches of the l2-regulariser rule."""
    from golden.replay import rel, replay_oracle
    n = 0
    for it, g, res, gm in replay_oracle(case):
        assert np.array_equal(res["mapping"], g[f"mapping{it}"])
        if f"samples{it}" in g.files:
            assert rel(res["samples"], g[f"samples{it}"]) < 1e-12
            assert rel(res["grads"], g[f"grads{it}"]) < 1e-10
        assert rel(res["bg"], g[f"bg{it}"]) < 1e-12
        assert rel(res["lnpdfs"], g[f"lnpdfs{it}"]) < 1e-12
        assert rel(res["H_neg"], g[f"H{it}"]) < 1e-9          # MORE: a 28-feature regression, 1e-13 measured
        assert rel(res["g_neg"], g[f"g{it}"]) < 1e-9
        assert rel(gm.means, g[f"means{it}"]) < 1e-10
        assert rel(gm.chol_cov, g[f"chol{it}"]) < 1e-10
        assert rel(gm.log_weights, g[f"log_weights{it}"]) < 1e-10
        assert rel(gm.stepsizes, g[f"stepsizes{it}"]) < 1e-12
        assert rel(gm.l2_regularizers, g[f"l2{it}"]) < 1e-12
        assert rel(gm.last_log_etas, g[f"last_log_etas{it}"]) < 1e-9
        assert np.array_equal(gm.num_received_updates, g[f"num_received_updates{it}"])
        if f"reward_history{it}" in g.files:
            for mine, ref in ((gm.reward_history, g[f"reward_history{it}"]), (gm.weight_history, g[f"weight_history{it}"])):
                unset = ref == -np.finfo(np.float32).max              # never written: tf.float32.min, exactly
                assert mine.shape == ref.shape and np.array_equal(mine == -np.finfo(np.float32).max, unset)
                assert np.allclose(mine[~unset], ref[~unset], rtol=1e-9, atol=1e-300)
        n += 1
    assert n == int(g["iterations"])


@pytest.mark.parametrize("diagonal", [False, True])
def test_oracle_model_surface_matches_reference_sources(diagonal):
    """The model API as the reference's FullCovGMM / DiagonalGMM compute it: densities, the
    GradientTape gradient, marginals, entropies, categorical sampling, GMM.sample's grouping (quirk 4), sampling with
    injected noise, weight replacement, adding and removing components."""
    from golden.replay import load, rel
    g = load("model_api_diagonal" if diagonal else "model_api_full")
    K, D = g["init_means"].shape
    chols, X = g["init_chols"], g["X"]
    dt = np.float64
    if diagonal:
        gm = O.make_diag_gmm(g["weights_in"], g["init_means"], np.stack([np.diag(c) ** 2 for c in chols]), dt)
    else:
        gm = O.make_full_gmm(g["weights_in"], g["init_means"], chols @ chols.transpose(0, 2, 1), dt)
    assert rel(gm.log_weights, g["log_weights"]) < 1e-13
    assert rel(O.component_log_densities(gm, X), g["component_log_densities"]) < 1e-13
    lq, lqk = O.log_densities_also_individual(gm, X)
    assert rel(lq, g["log_density"]) < 1e-13 and rel(lqk, g["individual"]) < 1e-13
    assert rel(O.log_density(gm, X), g["log_density_only"]) < 1e-13
    assert rel(np.exp(O.log_density(gm, X)), g["density"]) < 1e-12
    lq2, grad, lqk2 = O.log_density_and_grad(gm, X)
    assert rel(lq2, g["lq_grad"]) < 1e-13 and rel(grad, g["grad"]) < 1e-11 and rel(lqk2, g["lqk_grad"]) < 1e-13
    assert rel(O.component_entropies(gm), g["component_entropies"]) < 1e-13
    assert rel(O.get_average_entropy(gm), g["average_entropy"]) < 1e-13
    covs = np.stack([np.diag(c) ** 2 for c in chols]) if diagonal else chols @ chols.transpose(0, 2, 1)
    assert rel(covs, g["covs"]) < 1e-13
    if not diagonal:
        assert rel(O.component_log_densities(gm, X)[2], g["component_log_density_2"]) < 1e-13
        one = O.make_full_gmm(np.ones(1), g["init_means"][1:2], covs[1:2], dt)
        l1, g1, _ = O.log_density_and_grad(one, X)
        assert rel(l1, g["component_lq_1"]) < 1e-13 and rel(g1, g["component_grad_1"]) < 1e-11
        cm = O.component_marginal_log_densities(gm, X, 3)
        assert rel(cm, g["component_marginal_3"]) < 1e-13
        assert rel(O.logsumexp(cm + gm.log_weights[:, None], axis=0), g["marginal_3"]) < 1e-13
    # sampling
    comps = O.sample_categorical(gm, g["u"])
    assert np.array_equal(comps, g["sample_categorical"]) and comps[0] == K - 1      # u just below 1
    assert np.array_equal(g["sample_components"], comps)                               # draw order (quirk 4)
    counts = np.bincount(comps, minlength=K)
    assert np.array_equal(g["sample_noise_shapes"][:, 1], counts)
    offs = np.concatenate(([0], np.cumsum(counts)))
    xs, _ = O.sample_from_components_no_shuffle(gm, counts, lambda k, D_, n: g["sample_noise"][offs[k]:offs[k + 1]].T)
    assert rel(xs, g["sample_x"]) < 1e-13                                              # grouped by component
    n_per = g["n_per"]
    offs = np.concatenate(([0], np.cumsum(n_per)))
    xs2, mapping = O.sample_from_components_no_shuffle(gm, n_per,
                                                       lambda k, D_, n: g["no_shuffle_noise"][offs[k]:offs[k + 1]].T)
    assert np.array_equal(mapping, g["no_shuffle_mapping"]) and rel(xs2, g["no_shuffle_x"]) < 1e-13
    # structural edits
    rng = np.random.default_rng(78)                  # replay the generator's stream up to the replace_weights draw
    rng.dirichlet(np.ones(K)); rng.standard_normal(X.shape); rng.uniform(size=(30, 1))
    for s in g["sample_noise_shapes"]:
        rng.standard_normal(tuple(s))
    for k in range(K):
        rng.standard_normal((D, int(n_per[k])))
    new_lw = (np.log(g["weights_in"]) + rng.standard_normal(K)).astype(np.float32).astype(np.float64)
    gm.replace_weights(new_lw)
    assert rel(gm.log_weights, g["log_weights_replaced"]) < 1e-13
    O.add_component(gm, np.float32(0.2), (g["init_means"][0] + 1.0).astype(np.float32), g["new_cov"])
    assert rel(gm.log_weights, g["added_log_weights"]) < 1e-13 and rel(gm.chol_cov, g["added_chol"]) < 1e-13
    O.remove_component(gm, 1)
    assert rel(gm.log_weights, g["removed_log_weights"]) < 1e-13 and rel(gm.means, g["removed_means"]) < 1e-13


# ---------------------------------------------------------------------------------------------------------------------
# targets
# ---------------------------------------------------------------------------------------------------------------------
TARGET_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_targets.npz")


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-30))


@pytest.mark.parametrize("name", ["stm20", "stm_hard64"])
def test_student_t_mixture_oracle_matches_reference_class_and_scipy(name):
    """oracle.student_t_mixture_target against (a) the reference's StudentTMixture_LNPDF run over the tfp stand-in
    (tests/golden/make_reference_targets.py), (b) scipy.stats.multivariate_t, (c) central finite differences."""
    from scipy.stats import multivariate_t
    g = np.load(TARGET_GOLD)
    ch = g[name + "_chols"]
    K = ch.shape[0]
    covs = ch @ ch.transpose(0, 2, 1)
    w = np.ones(K) / K
    X = g[name + "_X"]
    f = O.student_t_mixture_target(w, g[name + "_means"], covs)
    v, gr = f(X)
    assert np.max(np.abs(v - g[name + "_lnpdf"])) < 1e-10 and rel(gr, g[name + "_grad"]) < 1e-11
    lp = np.stack([multivariate_t(loc=g[name + "_means"][j], shape=covs[j], df=2).logpdf(X) for j in range(K)])
    assert np.max(np.abs(O.logsumexp(lp + np.log(w)[:, None], axis=0) - v)) < 1e-9
    m = O.student_t_mixture_marginal_log_density(w, g[name + "_means"], covs, X, 3)
    assert np.max(np.abs(m - g[name + "_marg3"])) < 1e-12
    x0 = X[100:104]
    h = 1e-6
    for d in (0, 5):
        e = np.zeros(X.shape[1]); e[d] = h
        fd = (f(x0 + e)[0] - f(x0 - e)[0]) / (2 * h)
        assert np.allclose(fd, f(x0)[1][:, d], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,links,goals", [("planar10_4", 10, 4), ("planar10_1", 10, 1), ("planar3_4", 3, 4)])
def test_planar_robot_oracle_matches_reference_class(name, links, goals):
    g = np.load(TARGET_GOLD)
    X = g[name + "_X"]
    f = O.planar_robot_target(links, goals)
    v, gr = f(X)
    assert rel(v, g[name + "_lnpdf"]) < 1e-13 and rel(gr, g[name + "_grad"]) < 1e-13
    assert rel(O.planar_robot_forward_kinematics(X), g[name + "_fk"]) < 1e-14
    h = 1e-7
    e = np.zeros(links); e[1] = h
    fd = (f(X[:8] + e)[0] - f(X[:8] - e)[0]) / (2 * h)
    assert np.allclose(fd, gr[:8, 1], rtol=1e-4)
