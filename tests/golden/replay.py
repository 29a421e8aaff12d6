"""Replays the cases of tests/golden/reference_<case>.npz (outputs of the reference's own sources, see
make_reference_golden.py) with the oracle; shared by the CPU pin test and the GPU parity test."""
import os

import numpy as np

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
_IMPROVE = dict(min_stepsize=1e-4, max_stepsize=0.1, inc=1.1, dec=0.85)

# the oracle-side spelling of the configurations in make_reference_golden.CASES
CASES = {
    "samtron_fixed": dict(cfg=O.IterationConfig(desired_samples_per_component=60, weight_stepsize=0.05), stepsize=0.1),
    "samtron_d96": dict(cfg=O.IterationConfig(desired_samples_per_component=64, weight_stepsize=0.05), stepsize=0.1),
    "stein_standard_iw_direct": dict(cfg=O.IterationConfig(
        desired_samples_per_component=60, weight_stepsize=0.05, ng_self_normalized=False, updater="direct",
        weight_updater="direct", weight_self_normalized=False), stepsize=0.01),
    "stein_iblr_improvement": dict(cfg=O.IterationConfig(
        desired_samples_per_component=60, updater="iBLR", component_stepsize="improvement-based",
        component_stepsize_cfg=_IMPROVE), stepsize=0.01,
        weight_adapter=dict(initial=0.05, min_stepsize=1e-3, max_stepsize=1.0, inc=1.1, dec=0.85)),
    "more_trust_region": dict(cfg=O.IterationConfig(desired_samples_per_component=80, weight_stepsize=0.05,
                                                    ng_estimator="MORE"), stepsize=0.1, regularizer=1e-8),
    "diagonal_stein_trust_region": dict(cfg=O.IterationConfig(desired_samples_per_component=60, weight_stepsize=0.05),
                                        stepsize=0.1, diagonal=True),
    "samtron_adaptive": dict(cfg=O.IterationConfig(desired_samples_per_component=40, weight_stepsize=0.05), stepsize=0.1,
                             keep_samples=True,
                             adaptive=dict(del_iters=6, add_iters=2, max_components=9,
                                           thresholds_for_add_heuristic=[50.0, 10.0], min_weight_for_del_heuristic=0.05,
                                           num_database_samples=500)),
    "more_mixture_based": dict(cfg=O.IterationConfig(sample_selector="mixture-based", desired_samples_per_component=360,
                                                     ratio_reused_samples_to_desired=0.25, weight_stepsize=0.05,
                                                     ng_estimator="MORE"), stepsize=0.1, regularizer=1e-8,
                               keep_samples=True),
    "own_samples_decaying_temperature": dict(cfg=O.IterationConfig(
        desired_samples_per_component=60, only_use_own_samples=True, component_stepsize="decaying",
        component_stepsize_cfg=dict(initial_stepsize=0.1, annealing_exponent=0.5), weight_updater="direct",
        temperature=0.7), stepsize=0.1, weight_decay=dict(initial=0.1, annealing_exponent=0.5)),
    "single_component": dict(cfg=O.IterationConfig(desired_samples_per_component=60, weight_stepsize=0.05), stepsize=0.1),
    "direct_rejected_steps": dict(cfg=O.IterationConfig(desired_samples_per_component=60, weight_stepsize=0.05,
                                                        updater="direct"), stepsize=1.5),
    "samtron_reuse": dict(cfg=O.IterationConfig(desired_samples_per_component=60, weight_stepsize=0.05,
                                                ratio_reused_samples_to_desired=2.0), stepsize=0.1, keep_samples=True),
}


def load(name):
    return np.load(os.path.join(HERE, f"reference_{name}.npz"))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if a.size else 0.0


def replay_oracle(name, dt=np.float64):
    """Runs the oracle on the case's inputs and noise; yields (iteration, golden, oracle outputs, oracle model)."""
    g, oc = load(name), CASES[name]
    K, D = g["init_means"].shape
    kw = dict(initial_stepsize=oc["stepsize"])
    if "adaptive" in oc:                  # GmmWrapper.build_from_config, gmm_wrapper.py:53-54
        kw["max_reward_history_length"] = 2 * max(2, oc["adaptive"]["del_iters"])
    if "regularizer" in oc:
        kw["initial_regularizer"] = oc["regularizer"]
    chols, tchols = g["init_chols"], g["target_chols"]
    if oc.get("diagonal"):
        gm = O.make_diag_gmm(np.ones(K) / K, g["init_means"], np.stack([np.diag(c) ** 2 for c in chols]), dt, **kw)
    else:
        gm = O.make_full_gmm(np.ones(K) / K, g["init_means"], chols @ chols.transpose(0, 2, 1), dt, **kw)
    target = O.gmm_target(np.ones(3) / 3, g["target_means"], tchols @ tchols.transpose(0, 2, 1), dt)
    db = O.OracleSampleDB(D, bool(oc.get("diagonal")), bool(oc.get("keep_samples")),
                          100000 if oc.get("keep_samples") else None, dt)
    wad = O.ImprovementBasedWeightStepsize(dt=dt, **oc["weight_adapter"]) if "weight_adapter" in oc else None
    if "weight_decay" in oc:
        wad = O.DecayingWeightStepsize(dt=dt, **oc["weight_decay"])
    adapter = O.VipsComponentAdaptation(gm, db, 0.0, 1.0, **oc["adaptive"]) if "adaptive" in oc else None
    for it in range(int(g["iterations"])):
        noise, shapes = g[f"noise{it}"].astype(np.float64), g[f"noise_shapes{it}"]
        offs = np.concatenate(([0], np.cumsum(shapes[:, 1])))
        calls = [0]

        def noise_fn(k, D_, n):
            i = calls[0]
            calls[0] += 1
            assert tuple(shapes[i]) == (D_, n), f"draw {i}: the reference drew {tuple(shapes[i])}, the oracle asks {(D_, n)}"
            return noise[offs[i]:offs[i + 1]].T
        us = list(g[f"uniform{it}"]) if oc["cfg"].sample_selector == "mixture-based" else []

        def uniform_fn(n):
            assert n == len(us), f"the reference drew {len(us)} uniform numbers, the oracle asks {n}"
            return np.asarray(us)
        res = O.train_iter(gm, db, target, oc["cfg"], noise_fn, wad, uniform_fn)
        assert calls[0] == len(shapes), "number of sampling calls differs from the reference"
        if adapter is not None:           # GMMVI.train_iter: adapt_number_of_components(num_updates), gmmvi.py:161
            res["pre_adaptation"] = dict(means=gm.means.copy(), chol=gm.chol_cov.copy(), log_weights=gm.log_weights.copy())
            us, perm = list(g[f"uniform{it}"]), g[f"perm{it}"]

            def shuffle_fn(n):
                assert n == len(perm), (n, len(perm))
                return perm
            res["deleted"], res["added"] = adapter.adapt_number_of_components(it + 1, lambda: us.pop(0), shuffle_fn, target)
            assert not us, "the reference drew a uniform number the oracle did not ask for"
        yield it, g, res, gm
