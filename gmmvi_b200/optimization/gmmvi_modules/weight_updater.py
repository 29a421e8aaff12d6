"""Weight updaters (mirror of optimization/gmmvi_modules/weight_updater.py:5-281)."""
from __future__ import annotations

import torch

from ... import ops


class WeightUpdater:
    _trust_region = False

    def __init__(self, model, temperature: float, use_self_normalized_importance_weights: bool):
        self.model = model
        self.temperature = temperature
        self.use_self_normalized_importance_weights = use_self_normalized_importance_weights

    @staticmethod
    def build_from_config(config, gmm_wrapper):
        """weight_updater.py:33-54."""
        t = config["weight_updater_type"]
        if t == "direct":
            return DirectWeightUpdater(gmm_wrapper, temperature=config["temperature"], **config["weight_updater_config"])
        elif t == "trust-region":
            return TrustRegionBasedWeightUpdater(gmm_wrapper, temperature=config["temperature"], **config["weight_updater_config"])
        raise ValueError(f"config['weight_updater_type'] is '{t}' which is an unknown type")

    def _get_expected_log_ratios(self, samples, background_mixture_densities, target_lnpdfs):
        """weight_updater.py:56-75 (evaluated on the already-updated components, quirk 15)."""
        model_densities, lq = self.model.log_densities_also_individual(samples.contiguous())
        log_ratios = (target_lnpdfs - self.temperature * model_densities).contiguous()
        shard = self.model.shard
        if shard is not None:
            n_total = shard.all_reduce_sum_(torch.tensor([float(lq.shape[1])], device=lq.device)).item() \
                if not self.use_self_normalized_importance_weights else None
            elr = ops.importance_weights_sharded(lq, background_mixture_densities, shard,
                                                 self.use_self_normalized_importance_weights, rho=log_ratios,
                                                 want_dot=True, n_total=n_total)["dot"]
        else:
            elr = ops.importance_weights(lq, background_mixture_densities, None,
                                         self.use_self_normalized_importance_weights, rho=log_ratios,
                                         want_dot=True)["dot"]
        self.model.store_rewards(self.temperature * self.model.log_weights + elr)
        return elr

    def update_weights(self, samples, background_mixture_densities, target_lnpdfs, stepsize):
        """weight_updater.py:77-100."""
        elr = self._get_expected_log_ratios(samples, background_mixture_densities, target_lnpdfs)
        self._update_weights_from_expected_log_ratios(elr, stepsize)

    def _update_weights_from_expected_log_ratios(self, expected_log_ratios, stepsize):
        """weight_updater.py:123-141 (direct) / :262-279 (trust region); `stepsize` may be a device scalar."""
        if self.model.num_components > 1:
            new_lw, info = ops.weight_update(self._trust_region, self.model.log_weights, expected_log_ratios, stepsize,
                                             self.temperature)
            self.last_info = info
            self.model.replace_weights(new_lw)


class DirectWeightUpdater(WeightUpdater):
    _trust_region = False


class TrustRegionBasedWeightUpdater(WeightUpdater):
    _trust_region = True
