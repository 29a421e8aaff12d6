"""The configurations of the reference-source golden cases (tests/golden/reference_<case>.npz): plain data, importable
without TensorFlow, torch or /root/reference.  make_reference_golden.py runs the reference with them, tests/golden/replay.py
holds the oracle-side spelling, tests/test_api_gpu.py builds the device modules from the same dicts."""

K, D, DESIRED = 4, 6, 60          # defaults; a case may override (K, D)


def base_config(**over):
    cfg = {
        "temperature": 1.0, "use_sample_database": False, "max_database_size": 100000,
        "model_initialization": {"use_diagonal_covs": False, "prior_mean": 0.0, "initial_cov": 1.0},
        "ng_estimator_type": "Stein",
        "ng_estimator_config": {"only_use_own_samples": False, "use_self_normalized_importance_weights": True},
        "num_component_adapter_type": "fixed", "num_component_adapter_config": {},
        "sample_selector_type": "component-based",
        "sample_selector_config": {"desired_samples_per_component": DESIRED, "ratio_reused_samples_to_desired": 0.0},
        "ng_based_updater_type": "trust-region", "ng_based_updater_config": {},
        "component_stepsize_adapter_type": "fixed", "component_stepsize_adapter_config": {"initial_stepsize": 0.1},
        "weight_updater_type": "trust-region", "weight_updater_config": {"use_self_normalized_importance_weights": True},
        "weight_stepsize_adapter_type": "fixed", "weight_stepsize_adapter_config": {"initial_stepsize": 0.05},
    }
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(cfg.get(k), dict):
            cfg[k] = {**cfg[k], **v}
        else:
            cfg[k] = v
    return cfg


CASES = {
    # name: (config overrides, iterations, diagonal[, (K, D)])
    "samtron_fixed": ({}, 4, False),
    # one iteration at a size that takes the tensor-core kernels on the device (D = 96: fp16-split log-density with the
    # A operand in tensor memory, tensor-core Stein statistics and mixture gradient, blocked update)
    "samtron_d96": ({"sample_selector_config": {"desired_samples_per_component": 64}}, 1, False, (6, 96)),
    "stein_standard_iw_direct": ({"ng_estimator_config": {"use_self_normalized_importance_weights": False},
                                 "ng_based_updater_type": "direct", "weight_updater_type": "direct",
                                 "component_stepsize_adapter_config": {"initial_stepsize": 0.01},
                                 "weight_updater_config": {"use_self_normalized_importance_weights": False}}, 3, False),
    "stein_iblr_improvement": ({"ng_based_updater_type": "iBLR",
                                "component_stepsize_adapter_type": "improvement-based",
                                "component_stepsize_adapter_config": {"initial_stepsize": 0.01, "min_stepsize": 1e-4,
                                                                      "max_stepsize": 0.1, "stepsize_inc_factor": 1.1,
                                                                      "stepsize_dec_factor": 0.85},
                                "weight_stepsize_adapter_type": "improvement_based",
                                "weight_stepsize_adapter_config": {"initial_stepsize": 0.05, "min_stepsize": 1e-3,
                                                                   "max_stepsize": 1.0, "stepsize_inc_factor": 1.1,
                                                                   "stepsize_dec_factor": 0.85}}, 4, False),
    "more_trust_region": ({"ng_estimator_type": "MORE",
                           "ng_estimator_config": {"initial_l2_regularizer": 1e-8},
                           "sample_selector_config": {"desired_samples_per_component": 80}}, 3, False),
    "diagonal_stein_trust_region": ({"model_initialization": {"use_diagonal_covs": True}}, 3, True),
    # SAMTRON proper: components are added (every 2nd iteration, at the best of 500 database samples) and deleted
    # (Gaussian-smoothed reward window of del_iters = 6) by VipsComponentAdaptation
    "samtron_adaptive": ({"use_sample_database": True, "num_component_adapter_type": "adaptive",
                          "num_component_adapter_config": {"del_iters": 6, "add_iters": 2, "max_components": 9,
                                                           "thresholds_for_add_heuristic": [50.0, 10.0],
                                                           "min_weight_for_del_heuristic": 0.05,
                                                           "num_database_samples": 500, "num_prior_samples": 0},
                          "sample_selector_config": {"desired_samples_per_component": 40}}, 16, False),
    # the algorithm of BASELINE config C3: samples drawn from the mixture (LinSampleSelector, the number is the total),
    # MORE estimator, trust-region updates; with reuse of the newest database samples
    "more_mixture_based": ({"ng_estimator_type": "MORE", "ng_estimator_config": {"initial_l2_regularizer": 1e-8},
                            "use_sample_database": True, "sample_selector_type": "mixture-based",
                            "sample_selector_config": {"desired_samples_per_component": 360,
                                                       "ratio_reused_samples_to_desired": 0.25}}, 4, False),
    # quirk 6 (own samples only: uniform importance weights), decaying component and weight stepsizes, temperature != 1
    # (weight updater and eta = max(eta, temperature), but not the Stein log-ratios: quirk 7)
    "own_samples_decaying_temperature": ({"temperature": 0.7,
                                          "ng_estimator_config": {"only_use_own_samples": True},
                                          "component_stepsize_adapter_type": "decaying",
                                          "component_stepsize_adapter_config": {"initial_stepsize": 0.1,
                                                                                "annealing_exponent": 0.5},
                                          "weight_updater_type": "direct",
                                          "weight_stepsize_adapter_type": "decaying",
                                          "weight_stepsize_adapter_config": {"initial_stepsize": 0.1,
                                                                             "annealing_exponent": 0.5}}, 4, False),
    # one component: the weight updaters do nothing (quirk 13)
    "single_component": ({}, 2, False, (1, 6)),
    # direct updater with a step far too large: precision matrices lose positive definiteness, the failed Cholesky
    # rejects the step and the l2 regulariser follows the failure rule (quirk 11)
    "direct_rejected_steps": ({"ng_based_updater_type": "direct",
                               "component_stepsize_adapter_config": {"initial_stepsize": 1.5}}, 3, False),
    "samtron_reuse": ({"use_sample_database": True,
                       "sample_selector_config": {"ratio_reused_samples_to_desired": 2.0}}, 4, False),
}
