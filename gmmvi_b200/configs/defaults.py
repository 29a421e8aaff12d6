"""Default hyper-parameters: the values of the reference's configs/module_configs/**.yml (one per codeword
letter, configs/__init__.py:17-41) and of the experiment configs used by the BASELINE workloads."""

_IMPROVEMENT = dict(initial_stepsize=1.0, min_stepsize=0.001, max_stepsize=1.0,
                    stepsize_inc_factor=1.15, stepsize_dec_factor=0.85)

LETTER_TO_MODULE_CONFIG = {
    # NG estimator
    "Z": dict(ng_estimator_type="MORE",
              ng_estimator_config=dict(initial_l2_regularizer=1e-12, only_use_own_samples=False,
                                       use_self_normalized_importance_weights=True)),
    "S": dict(ng_estimator_type="Stein",
              ng_estimator_config=dict(only_use_own_samples=False, use_self_normalized_importance_weights=True)),
    # number of components
    "A": dict(num_component_adapter_type="adaptive",
              num_component_adapter_config=dict(del_iters=100, add_iters=30, max_components=1000,
                                                thresholds_for_add_heuristic=[5000.0, 1000.0, 500.0, 200.0, 100.0, 50.0],
                                                min_weight_for_del_heuristic=1.0e-6, num_database_samples=100000,
                                                num_prior_samples=0)),
    "E": dict(num_component_adapter_type="fixed", num_component_adapter_config=dict()),
    # sample selection
    "P": dict(sample_selector_type="mixture-based",
              sample_selector_config=dict(desired_samples_per_component=100, ratio_reused_samples_to_desired=0.0)),
    "M": dict(sample_selector_type="component-based",
              sample_selector_config=dict(desired_samples_per_component=100, ratio_reused_samples_to_desired=2.0)),
    # component update
    "I": dict(ng_based_updater_type="direct", ng_based_updater_config=dict()),
    "Y": dict(ng_based_updater_type="iBLR", ng_based_updater_config=dict()),
    "T": dict(ng_based_updater_type="trust-region", ng_based_updater_config=dict()),
    # component stepsize
    "F": dict(component_stepsize_adapter_type="fixed", component_stepsize_adapter_config=dict(initial_stepsize=1.0e-5)),
    "D": dict(component_stepsize_adapter_type="decaying",
              component_stepsize_adapter_config=dict(initial_stepsize=1.0, annealing_exponent=0.55)),
    "R": dict(component_stepsize_adapter_type="improvement-based",
              component_stepsize_adapter_config=dict(_IMPROVEMENT)),
    # weight update
    "U": dict(weight_updater_type="direct", weight_updater_config=dict(use_self_normalized_importance_weights=True)),
    "O": dict(weight_updater_type="trust-region",
              weight_updater_config=dict(use_self_normalized_importance_weights=True)),
    # weight stepsize
    "X": dict(weight_stepsize_adapter_type="fixed", weight_stepsize_adapter_config=dict(initial_stepsize=1.0)),
    "G": dict(weight_stepsize_adapter_type="decaying",
              weight_stepsize_adapter_config=dict(initial_stepsize=1.0, annealing_exponent="TODO")),  # sic (reference)
    "N": dict(weight_stepsize_adapter_type="improvement_based",
              weight_stepsize_adapter_config=dict(initial_stepsize=1.0, min_stepsize=0.0001, max_stepsize=1.0,
                                                  stepsize_inc_factor=1.15, stepsize_dec_factor=0.85)),
}


def _experiment(name, env_cfg, init, log_interval, max_db=10000000):
    return dict(start_seed=10000, environment_name=name, environment_config=env_cfg, model_initialization=init,
                gmmvi_runner_config=dict(log_metrics_interval=log_interval), use_sample_database=True,
                max_database_size=max_db, temperature=1.0)


EXPERIMENT_CONFIGS = {
    "stm20": _experiment("STM", dict(num_dimensions=20, harder_setting=False, use_matlab_target=False),
                         dict(use_diagonal_covs=False, num_initial_components=20, prior_mean=0.0, prior_scale=100.0,
                              initial_cov=300.0), 1000),
    "stm300": _experiment("STM", dict(num_dimensions=300, harder_setting=True, use_matlab_target=False),
                          dict(use_diagonal_covs=False, num_initial_components=20, prior_mean=0.0, prior_scale=100.0,
                               initial_cov=300.0), 50, 100000),
    "gmm20": _experiment("GMM", dict(num_dimensions=20),
                         dict(use_diagonal_covs=False, num_initial_components=1, prior_mean=0.0, prior_scale=31.63,
                              initial_cov=1000.0), 100),
    "gmm100": _experiment("GMM", dict(num_dimensions=100),
                          dict(use_diagonal_covs=False, num_initial_components=1, prior_mean=0.0, prior_scale=31.63,
                               initial_cov=1000.0), 100),
    "planar_robot_4": _experiment("PlanarRobot4", dict(),
                                  dict(use_diagonal_covs=False, num_initial_components=300, prior_mean=0.0,
                                       prior_scale=[1.0] + [0.2] * 9, initial_cov=[0.0625] + [0.0025] * 9), 10),
}
