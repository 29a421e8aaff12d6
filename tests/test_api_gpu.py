"""API-level parity: the module mirror (GMMVI / SampleSelector / NgEstimator / updaters) against the oracle's
`train_iter` on identical injected noise, plus runner smoke runs of the example configurations."""
import copy
import os

import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


# north_star tolerances: log-densities 1e-5, NG estimates and updated means / covariances 1e-4 (relative, fp32)
TOL_LOGDENS, TOL_NG = 1e-5, 1e-4
REPORT_ONLY = os.environ.get("GMMVI_B200_PARITY_REPORT", "0") == "1"     # print every measured error, assert nothing


def close(label, got, want, tol):
    """Assert rel_err(got, want) < tol; the measured error is printed (pytest -s) so that the bound can be audited."""
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    err = rel_err(got, want)
    print(f"PARITY {label}: {err:.3e} (tol {tol:.0e})")
    if not REPORT_ONLY:
        assert err < tol, f"{label}: {err:.3e} >= {tol:.0e}"


def base_config(updater="trust-region", weight_updater="trust-region", diag=False, desired=40, ratio=0.0,
                use_db=False, stepsize=0.05, comp_adapter="fixed"):
    return {
        "temperature": 1.0, "use_sample_database": use_db, "max_database_size": 100000,
        "model_initialization": {"use_diagonal_covs": diag, "prior_mean": 0.0, "initial_cov": 1.0},
        "ng_estimator_type": "Stein",
        "ng_estimator_config": {"only_use_own_samples": False, "use_self_normalized_importance_weights": True},
        "num_component_adapter_type": "fixed", "num_component_adapter_config": {},
        "sample_selector_type": "component-based",
        "sample_selector_config": {"desired_samples_per_component": desired, "ratio_reused_samples_to_desired": ratio},
        "ng_based_updater_type": updater, "ng_based_updater_config": {},
        "component_stepsize_adapter_type": comp_adapter,
        "component_stepsize_adapter_config": ({"initial_stepsize": stepsize} if comp_adapter == "fixed" else
                                              {"initial_stepsize": stepsize, "min_stepsize": 0.001, "max_stepsize": 1.0,
                                               "stepsize_inc_factor": 1.15, "stepsize_dec_factor": 0.85}),
        "weight_updater_type": weight_updater,
        "weight_updater_config": {"use_self_normalized_importance_weights": True},
        "weight_stepsize_adapter_type": "fixed", "weight_stepsize_adapter_config": {"initial_stepsize": 0.1},
    }


def build_pair(K, D, cfg, seed=0):
    from gmmvi_b200.models.full_cov_gmm import FullCovGMM
    from gmmvi_b200.models.diagonal_gmm import DiagonalGMM
    from gmmvi_b200.models.gmm_wrapper import GmmWrapper
    from gmmvi_b200.optimization.gmmvi import GMMVI
    from gmmvi_b200.experiments.target_distributions.gmm import GMM_LNPDF
    rng = np.random.default_rng(seed)
    diag = cfg["model_initialization"]["use_diagonal_covs"]
    means = (rng.standard_normal((K, D)) * 2).astype(np.float32)
    w = np.ones(K, np.float32) / K
    if diag:
        covs = rng.uniform(0.5, 2.0, (K, D)).astype(np.float32)
        model = DiagonalGMM(w, means, covs)
        og = O.make_diag_gmm(w, means, covs, np.float64)
    else:
        A = rng.standard_normal((K, D, D))
        covs = (A @ A.transpose(0, 2, 1) / D + np.eye(D)).astype(np.float32)
        model = FullCovGMM(w, means, covs)
        og = O.make_full_gmm(w, means, covs, np.float64)
    # keep the oracle on exactly the fp32 Cholesky factors the device holds
    og.chol_cov = model.chol_cov.cpu().numpy().astype(np.float64)
    tw = np.ones(3) / 3
    tm = rng.standard_normal((3, D)) * 2
    tA = rng.standard_normal((3, D, D))
    tc = tA @ tA.transpose(0, 2, 1) / D + np.eye(D)
    target = GMM_LNPDF(tw, tm, tc)
    otarget = O.gmm_target(tw, tm.astype(np.float32).astype(np.float64),
                           target.gmm.chol_cov.cpu().numpy().astype(np.float64) @
                           target.gmm.chol_cov.cpu().numpy().astype(np.float64).transpose(0, 2, 1), np.float64)
    wrapper = GmmWrapper.build_from_config(model, cfg)
    og.initial_stepsize = cfg["component_stepsize_adapter_config"]["initial_stepsize"]
    og.stepsizes = np.full(K, og.initial_stepsize)
    og.max_reward_history_length = 2
    gmmvi = GMMVI.build_from_config(cfg, target, wrapper)
    return gmmvi, og, otarget


@pytest.mark.parametrize("updater,weight_updater,diag", [("trust-region", "trust-region", False),
                                                         ("iBLR", "direct", False),
                                                         ("direct", "trust-region", False),
                                                         ("trust-region", "direct", True),
                                                         ("iBLR", "trust-region", True)])
def test_iterations_match_oracle(updater, weight_updater, diag):
    K, D, desired = 6, 12, 60
    step = 0.05 if updater != "direct" else 0.002
    cfg = base_config(updater, weight_updater, diag, desired, stepsize=step)
    gmmvi, og, otarget = build_pair(K, D, cfg)
    odb = O.OracleSampleDB(D, diag, False, None, np.float64)
    ocfg = O.IterationConfig(desired_samples_per_component=desired, updater=updater, weight_updater=weight_updater,
                             weight_stepsize=0.1)
    rng = np.random.default_rng(99)
    for it in range(3):
        E = rng.standard_normal((K * desired, D)).astype(np.float32)
        noise_fn = lambda k, D_, n: E[k * desired:(k + 1) * desired].T.astype(np.float64)
        out = O.train_iter(og, odb, otarget, ocfg, noise_fn)
        Ed = torch.as_tensor(E).cuda()
        samples, mapping, bg, lnpdfs, grads = gmmvi.sample_selector.select_samples(noise=Ed)
        assert np.array_equal(mapping.cpu().numpy(), out["mapping"])               # bit exact
        close("samples", samples, out["samples"], TOL_LOGDENS)
        close("bg", bg, out["bg"], TOL_LOGDENS)
        close("lnpdfs", lnpdfs, out["lnpdfs"], TOL_LOGDENS)
        close("grads", grads, out["grads"], TOL_NG)
        H, g = gmmvi.ng_estimator.get_expected_hessian_and_grad(samples, mapping, bg, lnpdfs, grads)
        close("H", H, out["H_neg"], TOL_NG)
        close("g", g, out["g_neg"], TOL_NG)
        gmmvi._run_updates(samples, mapping, bg, lnpdfs, grads)
        m = gmmvi.model
        assert np.array_equal(gmmvi.ng_based_updater.last_success.cpu().numpy().astype(bool), out["update"]["success"])
        close("means", m.means, og.means, TOL_NG)
        close("chol_cov", m.chol_cov, og.chol_cov, TOL_NG)
        close("weights", m.weights, og.weights, TOL_NG)
        assert np.allclose(m.l2_regularizers.cpu().numpy(), og.l2_regularizers)
        assert np.allclose(m.num_received_updates.cpu().numpy(), og.num_received_updates)
        if updater == "trust-region":
            assert np.allclose(m.last_log_etas.cpu().numpy(), og.last_log_etas, rtol=1e-4)


def test_sample_reuse_counts_are_exact():
    """n_add = max(1, desired - floor(ess)) must be bit exact (sample_selector.py:197-199)."""
    K, D, desired = 5, 8, 50
    cfg = base_config(desired=desired, ratio=2.0, use_db=True)
    gmmvi, og, otarget = build_pair(K, D, cfg, seed=3)
    odb = O.OracleSampleDB(D, False, True, 100000, np.float64)
    ocfg = O.IterationConfig(desired_samples_per_component=desired, ratio_reused_samples_to_desired=2.0,
                             weight_stepsize=0.1)
    rng = np.random.default_rng(5)
    for it in range(3):
        # the oracle decides how many samples are needed; both sides then consume the same noise rows
        old_bg, old_X, _, _, _ = odb.get_newest_samples(2 * desired * K)
        n_add = O.vips_num_additional_samples(og, old_X, old_bg, desired)
        tot = int(n_add.sum())
        E = rng.standard_normal((tot, D)).astype(np.float32)
        offs = np.concatenate(([0], np.cumsum(n_add)))
        noise_fn = lambda k, D_, n: E[offs[k]:offs[k + 1]].T.astype(np.float64)
        out = O.train_iter(og, odb, otarget, ocfg, noise_fn)
        samples, mapping, bg, lnpdfs, grads = gmmvi.sample_selector.select_samples(noise=torch.as_tensor(E).cuda())
        assert samples.shape[0] == out["samples"].shape[0], (it, n_add)
        assert np.array_equal(mapping.cpu().numpy(), out["mapping"])
        close("bg", bg, out["bg"], TOL_LOGDENS)
        gmmvi._run_updates(samples, mapping, bg, lnpdfs, grads)
        close("means", gmmvi.model.means, og.means, TOL_NG)


@pytest.mark.parametrize("experiment,codeword,overrides", [
    ("stm20", "SAMTRON", {"num_component_adapter_config": {"del_iters": 5, "add_iters": 2},
                          "sample_selector_config": {"desired_samples_per_component": 50, "ratio_reused_samples_to_desired": 0.0},
                          "model_initialization": {"num_initial_components": 8}}),
    ("planar_robot_4", "SAMTRON", {"num_component_adapter_config": {"del_iters": 5, "add_iters": 1},
                                   "sample_selector_config": {"desired_samples_per_component": 30, "ratio_reused_samples_to_desired": 0.0},
                                   "model_initialization": {"num_initial_components": 10}}),
    ("gmm20", "SEPYFUX", {"sample_selector_config": {"desired_samples_per_component": 200},
                          "component_stepsize_adapter_config": {"initial_stepsize": 0.001},
                          "model_initialization": {"num_initial_components": 5, "use_diagonal_covs": True}}),
    ("stm20", "SAMYROX", {"sample_selector_config": {"desired_samples_per_component": 40, "ratio_reused_samples_to_desired": 2.0},
                          "component_stepsize_adapter_config": {"initial_stepsize": 0.01, "max_stepsize": 0.1},
                          "model_initialization": {"num_initial_components": 6}}),
])
def test_runner_examples(experiment, codeword, overrides):
    """The example scripts' flow (examples/5_samtron_20D_student-T.py:13-31, 6_samtron_planar4.py:15-51)."""
    from gmmvi_b200.configs import get_default_algorithm_config, get_default_experiment_config, update_config
    from gmmvi_b200.gmmvi_runner import GmmviRunner
    algo = update_config(get_default_algorithm_config(codeword), overrides)
    env = update_config(get_default_experiment_config(experiment), {"start_seed": 1})
    config = update_config(env, algo)
    config["gmmvi_runner_config"] = {"log_metrics_interval": 6}
    runner = GmmviRunner.build_from_config(config)
    elbos = []
    for n in range(13):
        metrics = runner.iterate_and_log(n)
        assert set(["walltime", "num_samples", "num_components", "max_weight", "num_db_samples",
                    "num_db_components"]) <= set(metrics)
        if "-elbo" in metrics:
            elbos.append(-metrics["-elbo"])
    assert len(elbos) == 3 and all(np.isfinite(elbos))
    if codeword == "SAMTRON":
        assert elbos[-1] > elbos[0]    # the ELBO improves
    m = runner.gmmvi.model
    assert torch.isfinite(m.means).all() and torch.isfinite(m.chol_cov).all()
    assert abs(float(m.weights.sum().item()) - 1.0) < 1e-4


def test_unknown_types_raise():
    cfg = base_config()
    gmmvi, _, _ = build_pair(2, 3, cfg)
    from gmmvi_b200.optimization.gmmvi_modules.ng_estimator import NgEstimator
    from gmmvi_b200.optimization.gmmvi_modules.weight_updater import WeightUpdater
    bad = copy.deepcopy(cfg)
    bad["ng_estimator_type"] = "nope"
    with pytest.raises(ValueError):
        NgEstimator.build_from_config(bad, 1.0, gmmvi.model)
    bad["weight_updater_type"] = "nope"
    with pytest.raises(ValueError):
        WeightUpdater.build_from_config(bad, gmmvi.model)


def test_cpu_tensor_is_rejected():
    from gmmvi_b200 import ops, _lib
    with pytest.raises(_lib.GmmviLibraryError):
        ops.mixture_lse(torch.zeros(2, 3), torch.zeros(2))


@pytest.mark.parametrize("self_normalized,l2", [(True, 1e-8), (False, 1e-8), (False, 1e-6)])
def test_more_iteration_matches_oracle(self_normalized, l2):
    """MORE estimator + trust-region updates through the module API (BASELINE config C3's algorithm, small shape).
    Without self-normalisation the regression gets the plain exp(lw) (ng_estimator.py:352-356): the ridge term is not
    scale free, so weights exp(lw) / N (the Stein convention) would regularise N times harder -- visible at l2 = 1e-6,
    the value `l2_regularizers` reaches after failed updates."""
    K, D, desired = 3, 6, 400
    cfg = base_config("trust-region", "trust-region", False, desired, stepsize=0.05)
    cfg["ng_estimator_type"] = "MORE"
    cfg["ng_estimator_config"] = {"only_use_own_samples": False, "initial_l2_regularizer": l2,
                                  "use_self_normalized_importance_weights": self_normalized}
    cfg["ng_based_updater_config"] = {}
    gmmvi, og, otarget = build_pair(K, D, cfg, seed=4)
    gmmvi.model.initial_regularizer = l2
    og.initial_regularizer = l2
    og.l2_regularizers = np.full(K, l2)
    odb = O.OracleSampleDB(D, False, False, None, np.float64)
    ocfg = O.IterationConfig(desired_samples_per_component=desired, ng_estimator="MORE", weight_stepsize=0.1,
                             ng_self_normalized=self_normalized)
    rng = np.random.default_rng(123)
    for it in range(2):
        E = rng.standard_normal((K * desired, D)).astype(np.float32)
        out = O.train_iter(og, odb, otarget, ocfg, lambda k, D_, n: E[k * desired:(k + 1) * desired].T.astype(np.float64))
        samples, mapping, bg, lnpdfs, grads = gmmvi.sample_selector.select_samples(noise=torch.as_tensor(E).cuda())
        H, g = gmmvi.ng_estimator.get_expected_hessian_and_grad(samples, mapping, bg, lnpdfs, grads)
        close("H", H, out["H_neg"], TOL_NG)
        close("g", g, out["g_neg"], TOL_NG)
        gmmvi._run_updates(samples, mapping, bg, lnpdfs, grads)
        close("means", gmmvi.model.means, og.means, TOL_NG)
        close("chol_cov", gmmvi.model.chol_cov, og.chol_cov, TOL_NG)


@pytest.mark.parametrize("route", ["h16", "tf32", "simt"])
def test_more_c3_shape(route, monkeypatch):
    monkeypatch.setenv("GMMVI_B200_MORE_TC", {"h16": "1", "tf32": "tf32", "simt": "0"}[route])
    _more_c3_shape()


def _more_c3_shape():
    """BASELINE config C3 shape (D=100 -> F=5151 features): the blocked Cholesky path over 41 panels.  The normal
    matrix is only well conditioned in fp32 when N is a few times F (SURVEY.md "Hard parts": C3 with N=4096 < F is
    rank deficient in the reference too), hence N = 12000 here."""
    from gmmvi_b200 import ops
    from test_kernels_gpu import make_problem, gmm32_of, dev
    K, D, N = 1, 100, 12000
    g, X = make_problem(K, D, N, seed=9, scale=0.5)
    g32 = gmm32_of(g)
    g_in = O.OracleGMM(g32.log_weights.astype(np.float64), g32.means.astype(np.float64),
                       g32.chol_cov.astype(np.float64), False, initial_regularizer=1e-6)
    X64 = X.astype(np.float64)
    rng = np.random.default_rng(10)
    Q = rng.standard_normal((D, D)); Q = Q @ Q.T / D + np.eye(D)
    tl = (-0.5 * np.einsum("ni,ij,nj->n", X64, Q, X64)).astype(np.float32)
    mapping = np.sort(rng.integers(0, K, N)).astype(np.int32); mapping[-1] = K - 1
    lq64 = O.component_log_densities(g_in, X64)
    bg = O.logsumexp(lq64 + g_in.log_weights[:, None], axis=0).astype(np.float32)
    Href, gref = O.more_ng(g_in, X64, mapping, bg.astype(np.float64), tl.astype(np.float64))
    linv, prec, cst, _ = ops.prepare_full(dev(g32.chol_cov))
    lq = ops.logdens_full(dev(X), dev(g32.means), linv, cst)
    logq = ops.mixture_lse(lq, dev(g32.log_weights))
    iw = ops.importance_weights(lq, dev(bg), None, True, None, True)
    l2 = torch.full((K,), 1e-6, device="cuda")
    quad, lin, ok = ops.more_fit(l2, dev(X), dev(tl) - logq, iw["W"], dev(g32.means), linv)
    assert ok.cpu().numpy().all()
    # fp32 noise floor of this (ill-conditioned) regression: the oracle in fp32 mode, i.e. what the reference computes
    g_32 = O.OracleGMM(g32.log_weights, g32.means, g32.chol_cov, False, initial_regularizer=1e-6)
    H32, _ = O.more_ng(g_32, X, mapping, bg, tl)
    floor = rel_err(H32, Href)
    err = rel_err(quad.cpu().numpy(), Href)
    print(f"MORE C3 shape: device error {err:.3e}, fp32-oracle error {floor:.3e}")
    assert err < max(1e-3, 5 * floor), (err, floor)     # 5151-feature regression in fp32


@pytest.mark.parametrize("case", ["samtron_fixed", "samtron_d96", "diagonal_stein_trust_region", "stein_iblr_improvement",
                                  "own_samples_decaying_temperature", "single_component", "direct_rejected_steps",
                                  "stein_standard_iw_direct", "samtron_reuse"])
def test_first_iteration_matches_reference_sources(case):
    """The device path against outputs of the reference's OWN sources (tests/golden/reference_<case>.npz, produced by
    tests/golden/make_reference_golden.py from /root/reference/src/gmmvi in float64): one full iteration from the
    case's initial mixture on the case's noise.  All inputs of the fixtures are fp32-representable, so the device
    starts from exactly the reference's parameters.  samtron_d96 takes the tensor-core kernels (D = 96)."""
    from golden.cases import CASES, base_config as golden_config
    from golden.replay import load
    from gmmvi_b200.experiments.target_distributions.gmm import GMM_LNPDF
    from gmmvi_b200.models.diagonal_gmm import DiagonalGMM
    from gmmvi_b200.models.full_cov_gmm import FullCovGMM
    from gmmvi_b200.models.gmm_wrapper import GmmWrapper
    from gmmvi_b200.optimization.gmmvi import GMMVI
    g = load(case)
    over, _, diagonal = CASES[case][:3]
    cfg = golden_config(**over)
    K, D = g["init_means"].shape
    w = np.ones(K, np.float32) / K
    chols = g["init_chols"]
    if diagonal:
        model = DiagonalGMM(w, g["init_means"], np.stack([np.diag(c) ** 2 for c in chols]))
    else:
        model = FullCovGMM.from_cholesky(w, g["init_means"], chols)
    target = GMM_LNPDF.from_cholesky(np.ones(3) / 3, g["target_means"], g["target_chols"])
    gmmvi = GMMVI.build_from_config(cfg, target, GmmWrapper.build_from_config(model, cfg))
    noise = torch.as_tensor(np.asarray(g["noise0"], np.float32)).cuda()
    samples, mapping, bg, lnpdfs, grads = gmmvi.sample_selector.select_samples(noise=noise)
    assert np.array_equal(mapping.cpu().numpy(), g["mapping0"])                     # bit exact
    if "samples0" in g.files:
        close("samples", samples, g["samples0"], TOL_LOGDENS)
        close("grads", grads, g["grads0"], TOL_NG)
    close("bg", bg, g["bg0"], TOL_LOGDENS)
    close("lnpdfs", lnpdfs, g["lnpdfs0"], TOL_LOGDENS)
    H, gn = gmmvi.ng_estimator.get_expected_hessian_and_grad(samples, mapping, bg, lnpdfs, grads)
    close("H", H, g["H0"], TOL_NG)
    close("gn", gn, g["g0"], TOL_NG)
    gmmvi._run_updates(samples, mapping, bg, lnpdfs, grads)
    m = gmmvi.model
    chol_ref = np.stack([np.diag(c) for c in g["chol0"]]) if (diagonal and g["chol0"].ndim == 3) else g["chol0"]
    close("means", m.means, g["means0"], TOL_NG)
    close("chol_cov", m.chol_cov, chol_ref, TOL_NG)
    close("weights", m.weights, np.exp(g["log_weights0"]), TOL_NG)
    assert np.allclose(m.stepsizes.cpu().numpy(), g["stepsizes0"], rtol=1e-5)
    assert np.allclose(m.l2_regularizers.cpu().numpy(), g["l20"])
    assert np.array_equal(m.num_received_updates.cpu().numpy(), g["num_received_updates0"])
    if cfg["ng_based_updater_type"] == "trust-region":
        assert np.allclose(m.last_log_etas.cpu().numpy(), g["last_log_etas0"], rtol=1e-4)


def test_runner_dump_schema_and_metrics_match_oracle(tmp_path):
    """GmmviRunner.log_to_disk / finalize write the reference's .npz schema (gmmvi_runner.py:
    177-200: weights, means, covs, timestamps, fevals; `gmm_dump_<n>.npz` for n < 100 or n % 50 == 0, `final_gmm_dump.npz`),
    and the Monte-Carlo ELBO / entropy of get_expensive_metrics (:119-144) agree with the oracle on the same test samples."""
    import glob
    from gmmvi_b200.configs import get_default_algorithm_config, get_default_experiment_config, update_config
    from gmmvi_b200.gmmvi_runner import GmmviRunner
    algo = update_config(get_default_algorithm_config("SAMTRON"),
                         {"sample_selector_config": {"desired_samples_per_component": 40, "ratio_reused_samples_to_desired": 0.0},
                          "model_initialization": {"num_initial_components": 6}})
    config = update_config(update_config(get_default_experiment_config("gmm20"), {"start_seed": 3}), algo)
    config["gmmvi_runner_config"] = {"log_metrics_interval": 1}
    config["dump_gmm_path"] = str(tmp_path)
    runner = GmmviRunner.build_from_config(config)
    for n in range(3):
        out = runner.iterate_and_log(n)
        runner.log_to_disk(n)
    runner.log_to_disk(101)                      # neither < 100 nor a multiple of 50: nothing written
    runner.log_to_disk(150)
    runner.finalize()
    files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(runner.dump_gmm_path, "*.npz")))
    assert files == ["final_gmm_dump.npz", "gmm_dump_0.npz", "gmm_dump_1.npz", "gmm_dump_150.npz", "gmm_dump_2.npz"]
    m = runner.gmmvi.model
    K, D = m.num_components, m.num_dimensions
    d = np.load(os.path.join(runner.dump_gmm_path, "final_gmm_dump.npz"))
    assert sorted(d.files) == ["covs", "fevals", "means", "timestamps", "weights"]
    assert d["weights"].shape == (K,) and d["means"].shape == (K, D) and d["covs"].shape == (K, D, D)
    assert abs(float(d["weights"].sum()) - 1.0) < 1e-5 and int(d["fevals"]) == runner.gmmvi.sample_db.num_samples_written
    L = m.chol_cov.cpu().numpy().astype(np.float64)
    close("dumped covs", d["covs"], L @ L.transpose(0, 2, 1), TOL_LOGDENS)
    assert {"-elbo", "entropy", "target_density", "algo_time", "walltime", "num_samples", "num_components", "max_weight",
            "num_db_samples", "num_db_components"} <= set(out)
    # ELBO pieces on fixed test samples against the oracle (fp64) evaluated at the device's parameters
    test_samples, entropy = runner.get_samples_and_entropy(2000)
    og = O.OracleGMM(m.log_weights.cpu().numpy().astype(np.float64), m.means.cpu().numpy().astype(np.float64), L, False)
    X = test_samples.cpu().numpy().astype(np.float64)
    close("entropy", entropy.reshape(1), np.array([-np.mean(O.log_density(og, X))]), TOL_LOGDENS)
    tgt = runner.gmmvi.sample_selector.target_distribution
    tL = tgt.gmm.chol_cov.cpu().numpy().astype(np.float64)
    ot = O.OracleGMM(tgt.gmm.log_weights.cpu().numpy().astype(np.float64), tgt.gmm.means.cpu().numpy().astype(np.float64), tL, False)
    mean_reward = torch.mean(runner.gmmvi.sample_selector.target_uld(test_samples))
    close("target_density", mean_reward.reshape(1), np.array([np.mean(O.log_density(ot, X))]), TOL_LOGDENS)
