"""Experiment set-up (mirror of experiments/setup_experiment.py:10-160)."""
from __future__ import annotations

import numpy as np
import torch

from ..models.diagonal_gmm import DiagonalGMM
from ..models.full_cov_gmm import FullCovGMM
from ..models.gmm_wrapper import GmmWrapper


def init_experiment(config: dict, device="cuda"):
    """setup_experiment.py:10-43 -> (target, GmmWrapper with history length 10000, quirk 17)."""
    if "environment_config" in config.keys():
        target_fn = get_target_lnpdf(config["environment_name"], config["environment_config"], config["seed"], device)
    elif "target_fn" in config.keys():
        target_fn = config.pop("target_fn")
    else:
        raise ValueError("No target distribution was specified")
    gmm = construct_initial_mixture(num_dimensions=target_fn.get_num_dimensions(), device=device,
                                    **config["model_initialization"])
    initial_l2_regularizer = config["ng_estimator_config"].get("initial_l2_regularizer", 1e-12)
    gmm_wrapper = GmmWrapper(gmm, config["component_stepsize_adapter_config"]["initial_stepsize"],
                             initial_l2_regularizer, max_reward_history_length=10000)
    return target_fn, gmm_wrapper


def get_target_lnpdf(experiment, environment_config, seed, device="cuda"):
    """setup_experiment.py:46-86.  Targets outside the hot-path scope (logistic regression, BNN, Talos) are
    not provided; pass your own LNPDF through config['target_fn'] instead."""
    if experiment == "PlanarRobot4":
        from .target_distributions.planar_robot import make_four_goal
        return make_four_goal(device)
    elif experiment == "PlanarRobot1":
        from .target_distributions.planar_robot import make_single_goal
        return make_single_goal(device)
    elif experiment == "STM":
        from .target_distributions.student_t_mixture import make_target
        return make_target(device=device, **environment_config)
    elif experiment.startswith("GMM"):
        from .target_distributions.gmm import make_target
        return make_target(device=device, **environment_config)
    raise ValueError(f"get_target_lnpdf() was called with unknown experiment name: {experiment}")


def construct_initial_mixture(num_dimensions, num_initial_components, prior_mean, prior_scale, use_diagonal_covs,
                              initial_cov=None, device="cuda"):
    """setup_experiment.py:88-160: means ~ prior via the NumPy global RNG, covariances initial_cov * I."""
    if np.isscalar(prior_mean):
        prior_mean = prior_mean * np.ones(num_dimensions)
    if np.isscalar(prior_scale):
        prior_scale = prior_scale * np.ones(num_dimensions)
    prior = np.array(prior_scale) ** 2
    weights = np.ones(num_initial_components, dtype=np.float32) / num_initial_components
    means = np.zeros((num_initial_components, num_dimensions), dtype=np.float32)
    if use_diagonal_covs:
        initial_cov = prior if initial_cov is None else initial_cov * np.ones(num_dimensions)
        covs = np.ones((num_initial_components, num_dimensions), dtype=np.float32)
        for i in range(num_initial_components):
            if num_initial_components == 1:
                means[i] = prior_mean
            else:
                means[i] = prior_mean + np.sqrt(prior) * np.random.standard_normal([num_dimensions])
            covs[i] = initial_cov
        return DiagonalGMM(weights, means, covs, device=device)
    prior = np.diag(prior)
    initial_cov = prior if initial_cov is None else initial_cov * np.eye(num_dimensions)
    covs = np.ones((num_initial_components, num_dimensions, num_dimensions), dtype=np.float32)
    chol_prior = np.linalg.cholesky(prior)
    for i in range(num_initial_components):
        if num_initial_components == 1:
            means[i] = prior_mean
        else:
            means[i] = prior_mean + chol_prior @ np.random.standard_normal([num_dimensions, 1])[:, 0]
        covs[i] = initial_cov
    return FullCovGMM(weights, means, covs, device=device)
