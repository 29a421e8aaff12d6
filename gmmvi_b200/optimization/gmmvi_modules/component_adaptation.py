"""Adaptation of the number of components (mirror of optimization/gmmvi_modules/component_adaptation.py).
Host-side bookkeeping around the device log-density kernels."""
from __future__ import annotations

from math import pi, log, exp, sqrt

import torch

from ...models.diagonal_gmm import DiagonalGMM


class ComponentAdaptation:
    @staticmethod
    def build_from_config(config, gmm_wrapper, sample_db, target_distribution, prior_mean, initial_cov):
        """component_adaptation.py:48-82."""
        t = config["num_component_adapter_type"]
        if t == "adaptive":
            return VipsComponentAdaptation(gmm_wrapper, sample_db, target_distribution, prior_mean, initial_cov,
                                           **config["num_component_adapter_config"])
        elif t == "fixed":
            return FixedComponentAdaptation(**config["num_component_adapter_config"])
        raise ValueError(f"config['num_component_adapter_type'] is '{t}' which is an unknown type")

    def adapt_number_of_components(self, iteration, uniform=None, permutation=None):
        raise NotImplementedError


class FixedComponentAdaptation(ComponentAdaptation):
    def adapt_number_of_components(self, iteration, uniform=None, permutation=None):
        pass


class VipsComponentAdaptation(ComponentAdaptation):
    def __init__(self, model, sample_db, target_lnpdf, prior_mean, initial_cov, del_iters, add_iters, max_components,
                 thresholds_for_add_heuristic, min_weight_for_del_heuristic, num_database_samples, num_prior_samples):
        """component_adaptation.py:145-175."""
        self.model = model
        dev = model.device
        D = model.num_dimensions
        if prior_mean is not None and initial_cov is not None:
            pm = torch.as_tensor(prior_mean, dtype=torch.float32).reshape(-1)
            ic = torch.as_tensor(initial_cov, dtype=torch.float32).reshape(-1)
            if ic.numel() == 1:
                ic = ic * torch.ones(D)
            if pm.numel() == 1:
                pm = pm * torch.ones(D)
            self.prior = DiagonalGMM(torch.ones(1), pm.unsqueeze(0), ic.unsqueeze(0), device=dev)
        else:
            self.prior = None
        self.num_prior_samples = num_prior_samples
        self.target_lnpdf = target_lnpdf
        self.sample_db = sample_db
        self.del_iters = del_iters
        self.add_iters = add_iters
        self.max_components = max_components
        self.num_db_samples = num_database_samples
        self.num_calls_to_add_heuristic = 0
        th = thresholds_for_add_heuristic if isinstance(thresholds_for_add_heuristic, (list, tuple)) else [thresholds_for_add_heuristic]
        self.thresholds_for_addHeuristic = [float(t) for t in th]
        self.min_weight_for_del_heuristic = min_weight_for_del_heuristic
        self.reward_improvements = torch.zeros(0, device=dev)
        self.filter_delay = int(del_iters // 3)
        sigma = del_iters / 8.0
        xs = torch.arange(-self.filter_delay, self.filter_delay, dtype=torch.float32)
        kernel = torch.exp(-0.5 * (xs / sigma) ** 2) / (sigma * sqrt(2 * pi))
        self.kernel = (kernel / torch.sum(kernel)).to(dev)

    def adapt_number_of_components(self, iteration, uniform=None, permutation=None):
        """component_adaptation.py:177-190.  `uniform` / `permutation` (optional) replace the two random draws of an
        addition (the entropy mix of :208 and the database shuffle of sample_db.py:151) in the parity tests."""
        iteration = int(iteration)
        if iteration > self.del_iters:
            self.delete_bad_components()
        if iteration > 1 and iteration % self.add_iters == 0:
            if self.model.num_components < self.max_components:
                self.add_new_component(uniform, permutation)

    def add_at_best_location(self, samples, target_lnpdfs, uniform=None):
        """component_adaptation.py:193-226."""
        m = self.model
        D = m.num_dimensions
        it = self.num_calls_to_add_heuristic % len(self.thresholds_for_addHeuristic)
        thr = self.thresholds_for_addHeuristic[it]
        model_log_densities = m.log_density(samples.contiguous())
        init_weight = 1e-29
        a = torch.rand(1, device=m.device) if uniform is None else torch.full((1,), float(uniform), device=m.device)
        if self.prior is not None:
            des_entropy = m.get_average_entropy() * a + self.prior.get_average_entropy() * (1 - a)
        else:
            des_entropy = m.get_average_entropy().reshape(1)
        max_logdensity = torch.max(model_log_densities)
        rewards = target_lnpdfs - torch.maximum(max_logdensity - thr, model_log_densities)
        new_mean = samples[torch.argmax(rewards)]
        H_unscaled = 0.5 * D * (log(2.0 * pi) + 1)
        c = torch.exp((2 * (des_entropy - H_unscaled)) / D)
        if m.diagonal_covs:
            new_cov = c * torch.ones(D, device=m.device)
        else:
            new_cov = c * torch.eye(D, device=m.device)
        m.add_component(init_weight, new_mean, new_cov, torch.tensor([thr], device=m.device), des_entropy.reshape(1))

    def select_samples_for_adding_heuristic(self, permutation=None):
        """component_adaptation.py:229-249."""
        self.num_calls_to_add_heuristic += 1
        samples, target_lnpdfs = self.sample_db.get_random_sample(self.num_db_samples, permutation)
        prior_samples = torch.zeros((0, self.model.num_dimensions), device=self.model.device)
        if self.num_prior_samples > 0:
            prior_samples = self.prior.sample(self.num_prior_samples)[0]
            self.sample_db.num_samples_written += self.num_prior_samples
        return samples, target_lnpdfs, prior_samples

    def add_new_component(self, uniform=None, permutation=None):
        """component_adaptation.py:251-259."""
        samples, target_lnpdfs, prior_samples = self.select_samples_for_adding_heuristic(permutation)
        if self.num_prior_samples > 0:
            samples = torch.cat((samples, prior_samples), 0)
            target_lnpdfs = torch.cat((target_lnpdfs, self.target_lnpdf.log_density(prior_samples)), 0)
        self.add_at_best_location(samples, target_lnpdfs, uniform)

    def delete_bad_components(self):
        """component_adaptation.py:261-300."""
        m = self.model
        ks = self.kernel.numel()
        rh, wh = m.reward_history, m.weight_history
        kern = self.kernel.reshape(1, -1)
        current = torch.mean(rh[:, -ks:] * kern, dim=1)
        old = torch.mean(rh[:, -ks - self.del_iters:-self.del_iters] * kern, dim=1)
        old = old - torch.max(current)
        current = current - torch.max(current)
        reward_improvements = (current - old) / torch.abs(old)
        self.reward_improvements = reward_improvements
        max_actual_weights = torch.max(wh[:, -ks - self.del_iters:-1], dim=1).values
        window = rh[:, -ks - self.del_iters:]
        max_greedy_weights = torch.max(torch.exp(window - torch.logsumexp(window, dim=0, keepdim=True)), dim=1).values
        max_weights = torch.maximum(max_actual_weights, max_greedy_weights)
        is_stagnating = reward_improvements <= 0.4
        is_low_weight = max_weights < self.min_weight_for_del_heuristic
        is_old_enough = rh[:, -self.del_iters] != -torch.finfo(torch.float32).max
        is_bad = is_stagnating & is_low_weight & is_old_enough
        bad = torch.nonzero(is_bad).reshape(-1).tolist()
        for idx in sorted(bad, reverse=True):
            m.remove_component(idx)
