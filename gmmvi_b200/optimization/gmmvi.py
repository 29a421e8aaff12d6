"""GMMVI orchestrator (mirror of optimization/gmmvi.py:15-175): builds the seven modules from one config dict
and runs `select samples -> component update -> weight update -> adapt #components`."""
from __future__ import annotations

import torch

from .gmmvi_modules.component_adaptation import ComponentAdaptation
from .gmmvi_modules.component_stepsize_adaptation import ComponentStepsizeAdaptation
from .gmmvi_modules.ng_based_component_updater import NgBasedComponentUpdater
from .gmmvi_modules.ng_estimator import NgEstimator
from .gmmvi_modules.sample_selector import SampleSelector
from .gmmvi_modules.weight_stepsize_adaptation import WeightStepsizeAdaptation
from .gmmvi_modules.weight_updater import WeightUpdater
from .sample_db import SampleDB


class GMMVI:
    def __init__(self, model, sample_db, temperature, sample_selector, num_component_adapter,
                 component_stepsize_adapter, ng_estimator, ng_based_updater, weight_stepsize_adapter, weight_updater):
        """optimization/gmmvi.py:63-103 (no tf.function: the kernels are launched asynchronously on one stream)."""
        self.temperature = temperature
        self.model = model
        self.num_dimensions = model.num_dimensions
        self.sample_db = sample_db
        self.sample_selector = sample_selector
        self.num_component_adapter = num_component_adapter
        self.component_stepsize_adapter = component_stepsize_adapter
        self.ng_estimator = ng_estimator
        self.ng_based_updater = ng_based_updater
        self.weight_stepsize_adapter = weight_stepsize_adapter
        self.weight_updater = weight_updater
        self.num_updates = 0
        self._graph = None             # {'rng' | 'noise': optimization/graphed.GraphedIteration} when graphs are enabled
        self._graph_enabled = False
        self._graph_stable = 0         # eager iterations at the current number of components
        self._graph_patience = 1       # eager iterations required before a capture (backs off when K keeps changing)
        self._graph_pool = None        # memory pool shared by this object's graphs (optimization/graphed.py)
        self.graph_captures = 0

    @staticmethod
    def build_from_config(config: dict, target_distribution, model):
        """optimization/gmmvi.py:105-144."""
        sample_db = SampleDB.build_from_config(config, model.num_dimensions, device=model.device)
        ng_estimator = NgEstimator.build_from_config(config, config["temperature"], model)
        ng_based_updater = NgBasedComponentUpdater.build_from_config(config, model)
        num_component_adapter = ComponentAdaptation.build_from_config(
            config, model, sample_db, target_distribution=target_distribution,
            prior_mean=config["model_initialization"]["prior_mean"],
            initial_cov=config["model_initialization"]["initial_cov"])
        component_stepsize_adapter = ComponentStepsizeAdaptation.build_from_config(config, model)
        sample_selector = SampleSelector.build_from_config(config, model, sample_db, target_distribution)
        weight_updater = WeightUpdater.build_from_config(config, model)
        weight_stepsize_adapter = WeightStepsizeAdaptation.build_from_config(config, model)
        return GMMVI(model, sample_db, config["temperature"], sample_selector, num_component_adapter,
                     component_stepsize_adapter, ng_estimator, ng_based_updater, weight_stepsize_adapter,
                     weight_updater)

    def enable_sharding(self, shard):
        """Shard every iteration's samples (and the component update) over the ranks of `shard`
        (gmmvi_b200.distributed.ShardContext).  Supported for a fixed number of components without sample reuse."""
        from .gmmvi_modules.component_adaptation import FixedComponentAdaptation
        from .gmmvi_modules.sample_selector import VipsSampleSelector
        if not isinstance(self.num_component_adapter, FixedComponentAdaptation):
            raise NotImplementedError("multi-GPU sharding needs num_component_adapter_type='fixed'")
        if not isinstance(self.sample_selector, VipsSampleSelector) or self.sample_selector.reused_samples_per_component:
            raise NotImplementedError("multi-GPU sharding needs the component-based selector without sample reuse")
        if self.sample_db.keep_samples:
            raise NotImplementedError("multi-GPU sharding needs use_sample_database=False")
        gmm = self.model.model if hasattr(self.model, "model") else self.model
        gmm.shard = shard

    def enable_cuda_graph(self, enabled: bool = True):
        """Run the iteration as ONE captured CUDA graph (the counterpart of the reference's tf.function around train_iter,
        optimization/gmmvi.py:99-103): `train_iter()` replays the graph whenever the number of components has not changed
        since the previous iteration, and runs eagerly (then captures again) when it has.  Needs the component-based
        selector with ratio_reused_samples_to_desired = 0; results are bit-identical to eager iterations."""
        from .graphed import GraphedIteration
        if enabled:
            GraphedIteration(self)          # raises when the configuration cannot be captured
        self._graph_enabled, self._graph, self._graph_stable, self._graph_K = bool(enabled), None, 0, None
        if getattr(self, "_graph_state", None) is not None:
            self._graph_state.release()
        self._graph_state = None

    def _graphed_step(self, noise=None):
        """Replay (or capture, then replay) the graph of the current number of components -> True; False when this
        iteration has to run eagerly (first iteration at a new K).  Two graphs are kept: one that draws its noise from
        the device generator and one that reads injected noise from a static buffer the caller's tensor is copied to."""
        from .graphed import GraphedIteration
        K = self.model.num_components
        slot = "noise" if noise is not None else "rng"
        graphs = self._graph if isinstance(self._graph, dict) else {}
        if K != getattr(self, "_graph_K", None):
            # a capture costs tens of launches' worth of host time: if the number of components changes again before the
            # graph has paid for itself, wait longer before the next capture (examples/6 adds a component EVERY
            # iteration and ends up never capturing)
            if getattr(self, "_graph_K", None) is not None:
                replays = max((g.replays for g in graphs.values()), default=0)
                self._graph_patience = 1 if replays >= 16 else min(2 * self._graph_patience + 1, 63)
            # the retired graphs stay alive until the next capture has begun (the shared memory pool exists only while a
            # graph uses it) -- unless K keeps changing and the next capture is far away: then their memory is released
            self._graph_retired = ((list(graphs.values()) or getattr(self, "_graph_retired", None))
                                   if self._graph_patience <= 1 else None)
            graphs, self._graph_stable, self._graph_K = {}, 0, K
            if getattr(self, "_graph_state", None) is not None:
                self._graph_state.release()
                self._graph_state = None
        if noise is not None and slot in graphs and graphs[slot].noise_buffer.shape != noise.shape:
            del graphs[slot]
        self._graph = graphs
        if slot not in graphs:
            if self._graph_stable < self._graph_patience:      # first iteration(s) at this K run eagerly (warm-up)
                self._graph_stable += 1
                return False
            buf = None if noise is None else torch.empty_like(noise, memory_format=torch.contiguous_format)
            from .graphed import GraphCaptureError
            try:
                graphs[slot] = GraphedIteration(self, noise_buffer=buf).capture()
            except GraphCaptureError as e:
                # e.g. a user target that synchronises with the host: keep working, op by op
                import warnings
                warnings.warn(f"{e}; continuing with op-by-op iterations")
                self.enable_cuda_graph(False)
                return False
            self.graph_captures += 1
            self._graph_retired = None
        if noise is not None:
            graphs[slot].noise_buffer.copy_(noise)
        graphs[slot].replay()
        return True

    def train_iter(self, noise=None, adaptation_draws=None):
        """optimization/gmmvi.py:146-161.  `noise` ([N,D] standard-normal draws, optional) replaces the device
        generator for this iteration's samples (used by the parity tests and the end-to-end benchmark);
        `adaptation_draws` = (uniform, permutation) does the same for the two random draws of a component addition."""
        if not (self._graph_enabled and self._graphed_step(noise)):
            samples, mapping, sample_dist_densities, target_lnpdfs, target_lnpdf_grads = \
                self.sample_selector.select_samples(**({} if noise is None else {"noise": noise}))
            self._run_updates(samples, mapping, sample_dist_densities, target_lnpdfs, target_lnpdf_grads)
        if adaptation_draws is None:
            self.num_component_adapter.adapt_number_of_components(self.num_updates)
        else:
            self.num_component_adapter.adapt_number_of_components(self.num_updates, *adaptation_draws)

    def _run_updates(self, samples, mapping, sample_dist_densities, target_lnpdfs, target_lnpdf_grads):
        """optimization/gmmvi.py:163-174."""
        new_component_stepsizes = self.component_stepsize_adapter.update_stepsize(self.model.stepsizes)
        self.model.update_stepsizes(new_component_stepsizes)
        expected_hessian_neg, expected_grad_neg = self.ng_estimator.get_expected_hessian_and_grad(
            samples, mapping, sample_dist_densities, target_lnpdfs, target_lnpdf_grads)
        self.ng_based_updater.apply_NG_update(expected_hessian_neg, expected_grad_neg, self.model.stepsizes)
        weight_stepsize = self.weight_stepsize_adapter.update_stepsize()
        self.weight_updater.update_weights(samples, sample_dist_densities, target_lnpdfs, weight_stepsize)
        self.num_updates += 1
