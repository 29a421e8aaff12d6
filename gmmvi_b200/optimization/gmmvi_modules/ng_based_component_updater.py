"""Component updaters (mirror of optimization/gmmvi_modules/ng_based_component_updater.py:4-527)."""
from __future__ import annotations

import torch

from ... import ops


class NgBasedComponentUpdater:
    _mode = None

    def __init__(self, model, temperature: float):
        self.model = model
        self.temperature = temperature

    @staticmethod
    def build_from_config(config, gmm_wrapper):
        """ng_based_component_updater.py:31-57."""
        t = config["ng_based_updater_type"]
        kw = dict(temperature=config["temperature"], **config["ng_based_updater_config"])
        if t == "trust-region":
            return KLConstrainedNgBasedComponentUpdater(gmm_wrapper, **kw)
        elif t == "direct":
            return DirectNgBasedComponentUpdater(gmm_wrapper, **kw)
        elif t == "iBLR":
            return NgBasedComponentUpdaterIblr(gmm_wrapper, **kw)
        raise ValueError(f"config['ng_based_updater_type'] is '{t}' which is an unknown type")

    def apply_NG_update(self, expected_hessians_neg, expected_gradients_neg, stepsizes):
        """One batched kernel sequence updates every component; a component whose new precision is not positive
        definite keeps its old parameters (success flag instead of the reference's NaN test)."""
        m = self.model
        if self._mode == "direct" and m.diagonal_covs:
            raise NotImplementedError("the reference's direct updater has no diagonal-covariance branch "
                                      "(ng_based_component_updater.py:106 inverts a rank-1 tensor)")
        shard = m.shard
        rng_ = shard.component_range(m.num_components) if shard is not None else None
        general = (self._mode != "trust-region" and not m.diagonal_covs
                   and getattr(expected_hessians_neg, "gvi_nonsymmetric", False))
        if general:
            # non-symmetric -E[H] (Stein with standard importance weights, ng_estimator.py:168): the reference's general
            # inverse + Cholesky of its lower triangle, restated literally (every rank updates all components)
            _, prec, _ = m.prepared()
            means, chols, succ = ops.update_components_general(
                self._mode, m.means, m.chol_cov, prec, expected_hessians_neg, expected_gradients_neg, stepsizes,
                m.num_received_updates)
            etas = kls = None
            rng_ = None
        elif rng_ is None:
            means, chols, succ, etas, kls = ops.update_components(
                self._mode, m.diagonal_covs, m.means, m.chol_cov, expected_hessians_neg, expected_gradients_neg,
                stepsizes, m.last_log_etas, m.num_received_updates, self.temperature)
        else:
            # components sharded K/world per rank; the updated parameters are all-gathered
            a, b = rng_
            K = m.num_components
            sl = lambda t: t[a:b].contiguous()
            parts = ops.update_components(
                self._mode, m.diagonal_covs, sl(m.means), m.local_chol(a, b), sl(expected_hessians_neg),
                sl(expected_gradients_neg), sl(stepsizes), sl(m.last_log_etas), sl(m.num_received_updates),
                self.temperature)
            succ, etas, kls = m.set_components_sharded(a, b, parts[0], parts[1], (parts[2], parts[3], parts[4]))
            succ = succ.to(torch.int32)
        self.last_success, self.last_kls, self.last_etas = succ, kls, etas
        if rng_ is None:
            m.replace_components(means, chols)
        m.num_received_updates = m.num_received_updates + 1.0
        l2 = m.l2_regularizers                                            # quirk 11, :135-138
        m.l2_regularizers = torch.where(succ.bool(),
                                        torch.clamp(0.5 * l2, min=float(m.initial_regularizer)),
                                        torch.clamp(10.0 * l2, max=1e-6))
        if self._mode == "trust-region":
            m.last_log_etas = etas                                        # stores eta, not log eta (quirk 10)


class DirectNgBasedComponentUpdater(NgBasedComponentUpdater):
    _mode = "direct"


class NgBasedComponentUpdaterIblr(NgBasedComponentUpdater):
    _mode = "iBLR"


class KLConstrainedNgBasedComponentUpdater(NgBasedComponentUpdater):
    _mode = "trust-region"
