"""CUDA-graph iterations (gmmvi_b200/optimization/graphed.py, the counterpart of the reference's tf.function around
train_iter, optimization/gmmvi.py:99-103) reproduce eager iterations: same seed -> same noise subsequences -> the same
mixture, including across component additions / deletions (re-capture) and with the growing sample database."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fixed(K, D, desired, updater="trust-region", diag=False, variant=None):
    from gmmvi_b200.experiments.target_distributions.gmm import GMM_LNPDF
    from gmmvi_b200.models.diagonal_gmm import DiagonalGMM
    from gmmvi_b200.models.full_cov_gmm import FullCovGMM
    from gmmvi_b200.models.gmm_wrapper import GmmWrapper
    from gmmvi_b200.optimization.gmmvi import GMMVI
    from test_api_gpu import base_config
    rng = np.random.default_rng(0)
    means = (rng.standard_normal((K, D)) * 2).astype(np.float32)
    A = rng.standard_normal((K, D, D))
    covs = (A @ A.transpose(0, 2, 1) / D + np.eye(D)).astype(np.float32)
    tm = rng.standard_normal((3, D)) * 2
    tA = rng.standard_normal((3, D, D))
    cfg = base_config(updater, "trust-region", diag, desired, stepsize=0.05, comp_adapter="improvement-based")
    cfg["weight_stepsize_adapter_type"] = "improvement_based"
    cfg["weight_stepsize_adapter_config"] = {"initial_stepsize": 0.5, "min_stepsize": 0.0001, "max_stepsize": 1.0,
                                             "stepsize_inc_factor": 1.15, "stepsize_dec_factor": 0.85}
    if variant == "more":            # MORE estimator (ng_estimator.py:296-376) inside the graph
        cfg["ng_estimator_type"] = "MORE"
        cfg["ng_estimator_config"] = {"only_use_own_samples": False, "initial_l2_regularizer": 1e-8,
                                      "use_self_normalized_importance_weights": True}
    elif variant == "lin-stein":
        cfg["sample_selector_type"] = "mixture-based"
    elif variant == "lin-more":      # BASELINE C3's algorithm: mixture-based selector (desired = TOTAL samples) + MORE
        cfg["sample_selector_type"] = "mixture-based"
        cfg["ng_estimator_type"] = "MORE"
        cfg["ng_estimator_config"] = {"only_use_own_samples": False, "initial_l2_regularizer": 1e-8,
                                      "use_self_normalized_importance_weights": True}
    elif variant == "decaying":      # decaying stepsizes (device counters), temperature != 1, own samples only
        cfg["temperature"] = 0.7
        cfg["ng_estimator_config"]["only_use_own_samples"] = True
        cfg["component_stepsize_adapter_type"] = "decaying"
        cfg["component_stepsize_adapter_config"] = {"initial_stepsize": 0.05, "annealing_exponent": 0.6}
        cfg["weight_stepsize_adapter_type"] = "decaying"
        cfg["weight_stepsize_adapter_config"] = {"initial_stepsize": 0.5, "annealing_exponent": 0.4}
    w = np.ones(K, np.float32) / K
    model = DiagonalGMM(w, means, np.stack([np.diag(c) for c in covs])) if diag else FullCovGMM(w, means, covs)
    target = GMM_LNPDF(np.ones(3) / 3, tm, tA @ tA.transpose(0, 2, 1) / D + np.eye(D))
    return GMMVI.build_from_config(cfg, target, GmmWrapper.build_from_config(model, cfg))


def _state(g):
    m = g.model
    return {n: getattr(m, n).detach().cpu().numpy().copy() for n in
            ("means", "chol_cov", "log_weights", "stepsizes", "last_log_etas", "l2_regularizers", "num_received_updates",
             "reward_history", "weight_history")}


@pytest.mark.parametrize("K,D,desired,updater,diag,variant", [
    (8, 32, 64, "trust-region", False, None), (6, 96, 128, "trust-region", False, None), (5, 12, 50, "iBLR", False, None),
    (6, 16, 40, "trust-region", True, None), (3, 6, 400, "trust-region", False, "more"),
    (5, 20, 80, "trust-region", False, "decaying"), (4, 24, 60, "direct", False, "decaying"),
    (3, 6, 900, "trust-region", False, "lin-more"), (4, 40, 512, "trust-region", False, "lin-stein")])
def test_graph_replays_equal_eager_iterations(K, D, desired, updater, diag, variant):
    from gmmvi_b200 import rng
    iters = 6
    rng.set_seed(11)
    torch.manual_seed(123)           # the mixture-based selector draws its components from torch's generator
    eager = _fixed(K, D, desired, updater, diag, variant)
    for _ in range(iters):
        eager.train_iter()
    rng.set_seed(11)
    torch.manual_seed(123)
    graphed = _fixed(K, D, desired, updater, diag, variant)
    graphed.enable_cuda_graph()
    for _ in range(iters):
        graphed.train_iter()
    torch.cuda.synchronize()
    assert graphed._graph and graphed._graph["rng"].graph is not None           # iterations 2.. were replays
    assert graphed.num_updates == eager.num_updates == iters
    assert graphed.sample_db.num_samples_written == eager.sample_db.num_samples_written
    a, b = _state(eager), _state(graphed)
    for n in a:
        assert np.array_equal(a[n], b[n]), (n, float(np.max(np.abs(a[n] - b[n]))))
    # the database holds the last iteration's samples in both modes
    assert torch.equal(eager.sample_db.samples, graphed.sample_db.samples)
    assert torch.equal(eager.sample_db.mapping, graphed.sample_db.mapping)


def test_graph_with_injected_noise_equals_eager():
    """train_iter(noise=...) in graph mode: the noise is copied into the graph's static buffer before the replay."""
    K, D, desired, iters = 8, 32, 64, 5
    gen = torch.Generator(device="cuda").manual_seed(3)
    noise = [torch.randn((K * desired, D), device="cuda", generator=gen) for _ in range(iters)]
    eager, graphed = _fixed(K, D, desired), _fixed(K, D, desired)
    graphed.enable_cuda_graph()
    for e in noise:
        eager.train_iter(noise=e)
        graphed.train_iter(noise=e.clone())
    torch.cuda.synchronize()
    assert "noise" in graphed._graph
    a, b = _state(eager), _state(graphed)
    for n in a:
        assert np.array_equal(a[n], b[n]), (n, float(np.max(np.abs(a[n] - b[n]))))


def test_graph_mode_follows_component_adaptation():
    """SAMTRON with VipsComponentAdaptation and the growing sample database (the runner flow of examples/5): components
    are added / deleted between replays, the graph is captured again for every new K."""
    from gmmvi_b200.configs import get_default_algorithm_config, get_default_experiment_config, update_config
    from gmmvi_b200.gmmvi_runner import GmmviRunner

    def run(graph):
        algo = update_config(get_default_algorithm_config("SAMTRON"),
                             {"num_component_adapter_config": {"del_iters": 5, "add_iters": 4},
                              "sample_selector_config": {"desired_samples_per_component": 50, "ratio_reused_samples_to_desired": 0.0},
                              "model_initialization": {"num_initial_components": 8}})
        config = update_config(update_config(get_default_experiment_config("stm20"), {"start_seed": 5}), algo)
        config["gmmvi_runner_config"] = {"log_metrics_interval": 10 ** 9, "use_cuda_graph": graph}
        runner = GmmviRunner.build_from_config(config)
        assert runner.gmmvi._graph_enabled == graph
        ks = []
        for n in range(18):
            runner.iterate_and_log(n)
            ks.append(runner.gmmvi.model.num_components)
        return runner, ks
    torch.manual_seed(0)
    e, ks_e = run(False)
    torch.manual_seed(0)
    g, ks_g = run(True)
    assert ks_e == ks_g and len(set(ks_e)) > 1            # components were added (and the graph re-captured)
    assert e.gmmvi.sample_db.num_samples_written == g.gmmvi.sample_db.num_samples_written
    assert e.gmmvi.sample_db.samples.shape == g.gmmvi.sample_db.samples.shape
    a, b = _state(e.gmmvi), _state(g.gmmvi)
    for n in a:
        assert np.array_equal(a[n], b[n]), (n, float(np.max(np.abs(a[n] - b[n]))))
    assert torch.equal(e.gmmvi.sample_db.samples, g.gmmvi.sample_db.samples)
    assert torch.equal(e.gmmvi.sample_db.mapping, g.gmmvi.sample_db.mapping)


def test_rng_and_noise_graphs_share_their_state():
    """Alternating train_iter() and train_iter(noise=...) replays two different graphs over ONE set of static buffers."""
    from gmmvi_b200 import rng
    K, D, desired, iters = 6, 96, 128, 6
    gen = torch.Generator(device="cuda").manual_seed(4)
    noise = [torch.randn((K * desired, D), device="cuda", generator=gen) for _ in range(iters)]
    runs = []
    for graph in (False, True):
        rng.set_seed(21)
        g = _fixed(K, D, desired)
        if graph:
            g.enable_cuda_graph()
        for i in range(iters):
            g.train_iter()
            g.train_iter(noise=noise[i])
            if i == 2:                                  # eager code in between: an extra draw and a density evaluation
                X, _ = g.model.sample(100)
                g.model.log_density(X)
        torch.cuda.synchronize()
        runs.append(_state(g))
        if graph:
            assert set(g._graph) == {"rng", "noise"} and g._graph["rng"].replays >= iters - 2
    for n in runs[0]:
        assert np.array_equal(runs[0][n], runs[1][n]), (n, float(np.max(np.abs(runs[0][n] - runs[1][n]))))


def test_capture_failure_falls_back_to_eager_iterations():
    """A target that synchronises with the host inside log_density cannot be captured: graph mode must turn itself off,
    warn, and the run must continue op by op with the same results as a run that never tried."""
    from gmmvi_b200 import rng
    from gmmvi_b200.experiments.target_distributions.lnpdf import LNPDF

    class SyncingTarget(LNPDF):
        def __init__(self, inner):
            super().__init__(use_log_density_and_grad=True)
            self.inner = inner

        def get_num_dimensions(self):
            return self.inner.get_num_dimensions()

        def log_density_and_grad(self, x):
            v, g = self.inner.log_density_and_grad(x)
            float(v[0].item())                      # device -> host read: illegal while a stream is capturing
            return v, g

    results = []
    for graph in (False, True):
        rng.set_seed(5)
        g = _fixed(6, 32, 64)
        g.sample_selector.target_distribution = SyncingTarget(g.sample_selector.target_distribution)
        if graph:
            g.enable_cuda_graph()
            with pytest.warns(UserWarning, match="capture of the iteration failed"):
                for _ in range(4):
                    g.train_iter()
            assert not g._graph_enabled
            torch.randn(4, device="cuda")           # torch's CUDA generator was put back into its normal state
        else:
            for _ in range(4):
                g.train_iter()
        torch.cuda.synchronize()
        assert g.num_updates == 4
        results.append(_state(g))
    for n in results[0]:
        assert np.array_equal(results[0][n], results[1][n]), n


def test_eager_parameter_change_between_replays_is_picked_up():
    """User code that replaces the components between two graph replays (here: GMM.replace_components, models/gmm.py:
    401-418) must see the next replay start from the new parameters, their inverse factors and their split operands."""
    from gmmvi_b200 import rng
    K, D, desired = 6, 96, 128
    runs = []
    for graph in (False, True):
        rng.set_seed(31)
        g = _fixed(K, D, desired)
        if graph:
            g.enable_cuda_graph()
        for i in range(6):
            g.train_iter()
            if i == 3:
                m = g.model
                m.replace_components(m.means * 1.01 + 0.05, m.chol_cov * 1.1)
                m.log_density(torch.zeros((4, D), device="cuda"))        # forces the derived operands, eagerly
            if i == 4:                                                    # only O(K) state replaced
                g.model.replace_weights(g.model.log_weights + torch.linspace(0, 0.3, K, device="cuda"))
                g.model.update_stepsizes(g.model.stepsizes * 0.9)
        torch.cuda.synchronize()
        runs.append(_state(g))
    for n in runs[0]:
        assert np.array_equal(runs[0][n], runs[1][n]), (n, float(np.max(np.abs(runs[0][n] - runs[1][n]))))
