"""Natural-gradient estimators (mirror of optimization/gmmvi_modules/ng_estimator.py:10-376)."""
from __future__ import annotations

import torch

from ... import ops


class NgEstimator:
    def __init__(self, temperature, model, requires_gradient, only_use_own_samples, use_self_normalized_importance_weights):
        self._model = model
        self._temperature = temperature
        self._requires_gradients = requires_gradient
        self._only_use_own_samples = only_use_own_samples
        self._use_self_normalized_importance_weights = use_self_normalized_importance_weights

    @staticmethod
    def build_from_config(config, temperature, gmm_wrapper):
        """ng_estimator.py:47-65."""
        if config["ng_estimator_type"] == "Stein":
            return SteinNgEstimator(temperature=temperature, model=gmm_wrapper, **config["ng_estimator_config"])
        elif config["ng_estimator_type"] == "MORE":
            return MoreNgEstimator(temperature=temperature, model=gmm_wrapper, **config["ng_estimator_config"])
        raise ValueError(f"config['ng_estimator_type'] is '{config['ng_estimator_type']}' which is an unknown type")

    @property
    def requires_gradients(self) -> bool:
        return self._requires_gradients

    def get_expected_hessian_and_grad(self, samples, mapping, background_densities, target_lnpdfs, target_lnpdfs_grads):
        raise NotImplementedError

    def _relative_mapping(self, mapping):
        """ng_estimator.py:244,343 (quirk 3)."""
        return (mapping - torch.max(mapping) + self._model.num_components - 1).to(torch.int32).contiguous()

    def _importance_weights(self, lq, mapping, background_densities, unnormalized_mean=True, **want):
        """ng_estimator.py:107-120 + :173-176 / :155: W[K,N] per component, all on device.  Without self-normalisation
        the Stein estimator averages exp(lw) over the samples (weights exp(lw) / N, :147-155) while MORE hands the plain
        exp(lw) to its ridge regression (:352-356, `unnormalized_mean=False`): the ridge term is not scale free."""
        shard = self._model.shard
        if shard is not None:
            if self._only_use_own_samples:
                raise NotImplementedError("only_use_own_samples is not supported together with multi-GPU sharding")
            n_total = shard.all_reduce_sum_(torch.tensor([float(lq.shape[1])], device=lq.device)).item() \
                if not self._use_self_normalized_importance_weights else None
            return ops.importance_weights_sharded(lq, background_densities, shard,
                                                  self._use_self_normalized_importance_weights, n_total=n_total, **want)
        if self._only_use_own_samples:
            return ops.importance_weights(lq, None, self._relative_mapping(mapping), True, **want)
        mode = 1 if self._use_self_normalized_importance_weights else (0 if unnormalized_mean else 2)
        return ops.importance_weights(lq, background_densities, None, mode, **want)


class SteinNgEstimator(NgEstimator):
    def __init__(self, temperature, model, only_use_own_samples, use_self_normalized_importance_weights):
        super().__init__(temperature, model, True, only_use_own_samples, use_self_normalized_importance_weights)

    def get_expected_hessian_and_grad(self, samples, mapping, background_densities, target_lnpdfs, target_lnpdfs_grads):
        """ng_estimator.py:204-263 -> (expected_hessian_neg [K,D,D] | [K,D], expected_gradient_neg [K,D]).
        The reference loops over components with a cholesky_solve and an [N,D,D] outer product each; here the
        per-component sums are two batched contractions  M_k = sum_n w_kn (x_n-mu_k) g_n^T,  H_k = P_k M_k."""
        model = self._model
        samples = samples.contiguous()
        _, model_densities_grad, lq = model.log_density_and_grad(samples)
        G = (target_lnpdfs_grads - model_densities_grad).contiguous()
        iw = self._importance_weights(lq, mapping, background_densities, want_W=True, want_active=True)
        shard = model.shard
        rng_ = shard.component_range(model.num_components) if (shard is not None and shard.world > 1) else None
        symmetrize = self._use_self_normalized_importance_weights       # quirk 7
        if not symmetrize:
            rng_ = None         # the general (non-symmetric) direct / iBLR update is replicated: all rows must be valid
        self.valid_rows = None
        if model.diagonal_covs:
            H, g = ops.stein_diag(samples, model.means, model.chol_cov, iw["W"], G)
        elif rng_ is not None:
            # Sample-sharded run: the RAW statistics (sums over this rank's samples) are reduce-scattered by component
            # (half the traffic of an all-reduce) and only the K / world components this rank updates are finalised
            # (H_k = -sym(P_k M_k)).  Rows outside [a, b) of the returned tensors are never read (`valid_rows`).
            a, b = rng_
            M, g = ops.stein_stats_full(samples, model.means, iw["W"], iw["active"], G)
            M_own = shard.reduce_scatter_rows(M)
            g[a:b] = shard.reduce_scatter_rows(g)
            _, prec, _ = model.prepared()
            H = M                                                        # storage reuse: only rows [a, b) are valid
            H[a:b] = ops.stein_finalize_full(prec[a:b].contiguous(), M_own, symmetrize)
            self.valid_rows = (a, b)
            H.gvi_nonsymmetric = not symmetrize
            return H, g
        else:
            _, prec, _ = model.prepared()
            H, g = ops.stein_full(samples, model.means, prec, iw["W"], iw["active"], G, symmetrize)
            # the standard-IW estimate is not symmetric: the direct / iBLR updaters must treat it like the reference's
            # LU-based tf.linalg.inv / solve does (ng_based_component_updater.py:116-117, 199)
            H.gvi_nonsymmetric = not symmetrize
        if shard is not None and shard.world > 1:     # K not divisible by the world size: plain all-reduce
            nonsym = getattr(H, "gvi_nonsymmetric", False)
            shard.all_reduce_sum_(H)
            shard.all_reduce_sum_(g)
            H.gvi_nonsymmetric = nonsym
        return H, g


class MoreNgEstimator(NgEstimator):
    def __init__(self, temperature, model, only_use_own_samples, initial_l2_regularizer, use_self_normalized_importance_weights):
        super().__init__(temperature, model, True, only_use_own_samples, use_self_normalized_importance_weights)
        if not torch.all(model.l2_regularizers == float(initial_l2_regularizer) * torch.ones_like(model.l2_regularizers)):
            raise AssertionError("model.l2_regularizers != initial_l2_regularizer")      # ng_estimator.py:293
        from ..least_squares import QuadFunc
        self.least_square_fitter = QuadFunc(model.num_dimensions)

    def get_expected_hessian_and_grad(self, samples, mapping, background_densities, target_lnpdfs, target_lnpdfs_grads):
        """ng_estimator.py:296-376."""
        model = self._model
        samples = samples.contiguous()
        model_densities, lq = model.log_densities_also_individual(samples)
        log_ratios = (target_lnpdfs - model_densities).contiguous()
        iw = self._importance_weights(lq, mapping, background_densities, unnormalized_mean=False, want_W=True)
        if model.diagonal_covs:
            raise NotImplementedError("MORE does not support diagonal covariances (least_squares.py:172 inverts the "
                                      "Cholesky factor as a matrix)")
        if model.shard is not None:
            raise NotImplementedError("MORE is not supported together with multi-GPU sharding")
        linv, _, _ = model.prepared(need_prec=False)
        quad, lin, ok = self.least_square_fitter.fit_quadratic_batched(model.l2_regularizers, samples, log_ratios,
                                                                       iw["W"], model.means, linv)
        self.last_ok = ok
        # expected_gradient_neg = reward_quad mu - reward_lin   (ng_estimator.py:369-373)
        g = ops.bgemm(quad, model.means.unsqueeze(2).contiguous()).squeeze(2) - lin
        return quad, g.contiguous()
