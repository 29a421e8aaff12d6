"""Sample selection (mirror of optimization/gmmvi_modules/sample_selector.py:6-340)."""
from __future__ import annotations

import math

import torch

from ... import ops


class SampleSelector:
    def __init__(self, target_distribution, model, sample_db):
        self.target_distribution = target_distribution
        self.model = model
        self.sample_db = sample_db

    @staticmethod
    def build_from_config(config, gmm_wrapper, sample_db, target_distribution):
        """sample_selector.py:39-64."""
        if config["sample_selector_type"] == "component-based":
            return VipsSampleSelector(target_distribution, gmm_wrapper, sample_db, **config["sample_selector_config"])
        elif config["sample_selector_type"] == "mixture-based":
            return LinSampleSelector(target_distribution, gmm_wrapper, sample_db, **config["sample_selector_config"])
        raise ValueError(f"config['sample_selector_type'] is '{config['sample_selector_type']}' which is an unknown type")

    def target_uld(self, samples):
        return self.target_distribution.log_density(samples)

    def get_target_grads(self, samples):
        """sample_selector.py:69-78 -> (gradient, target)."""
        if self.target_distribution.use_log_density_and_grad:
            target, gradient = self.target_distribution.log_density_and_grad(samples)
        else:
            x = samples.detach().clone().requires_grad_(True)
            with torch.enable_grad():
                target = self.target_distribution.log_density(x)
                gradient, = torch.autograd.grad(target.sum(), x)
            target = target.detach()
        return gradient.contiguous(), target.contiguous()

    def select_samples(self):
        raise NotImplementedError

    def _prepared(self):
        return None if self.model.diagonal_covs else self.model.prepared(need_prec=False)


class VipsSampleSelector(SampleSelector):
    def __init__(self, target_distribution, model, sample_db, desired_samples_per_component, ratio_reused_samples_to_desired):
        super().__init__(target_distribution, model, sample_db)
        self.desired_samples_per_component = int(desired_samples_per_component)
        self.reused_samples_per_component = int(math.floor(ratio_reused_samples_to_desired * desired_samples_per_component))

    def get_effective_samples(self, model_densities, oldsamples_pdf):
        """sample_selector.py:140-158 -> ess[K] (float)."""
        return ops.importance_weights(model_densities, oldsamples_pdf, want_ess=True)["ess"]

    def sample_where_needed(self, samples, oldsamples_pdf, num_desired_samples=None, noise=None):
        """sample_selector.py:160-202 -> (new_samples, new_target_lnpdfs, new_target_grads, mapping)."""
        if num_desired_samples is None:
            num_desired_samples = self.desired_samples_per_component
        K = self.model.num_components
        shard = self.model.shard
        row_offset = 0
        own = None
        if shard is not None:
            # sample-sharded iteration (no reuse): this rank draws only its contiguous range of the global rows
            if samples.shape[0] != 0:
                raise NotImplementedError("sample reuse is not supported together with multi-GPU sharding")
            per = max(1, int(num_desired_samples))
            local, row_offset = shard.local_counts([per] * K)
            key = (tuple(local), str(self.model.device))       # constant across iterations: one upload, graph-safe
            if getattr(self, "_n_add_key", None) != key:
                self._n_add_key, self._n_add_dev = key, torch.tensor(local, device=self.model.device, dtype=torch.int32)
            n_add = self._n_add_dev
            total, mx = sum(local), max(local)
            rng_ = shard.component_range(K)
            if rng_ is not None and all(c == 0 for i, c in enumerate(local) if not rng_[0] <= i < rng_[1]):
                own = rng_           # every row of this rank belongs to a component it updates itself
            self.sample_db.count_override = torch.full((K,), float(per), device=self.model.device)
        elif samples.shape[0] == 0:
            n_add = torch.full((K,), max(1, int(num_desired_samples)), device=self.model.device, dtype=torch.int32)
            total, mx = K * max(1, int(num_desired_samples)), max(1, int(num_desired_samples))
        else:
            model_logpdfs = self.model.component_log_densities(samples)
            n_eff = torch.floor(self.get_effective_samples(model_logpdfs, oldsamples_pdf)).to(torch.int32)
            n_add = torch.clamp(num_desired_samples - n_eff, min=1).to(torch.int32)
            total, mx = None, None
        new_samples, mapping = self.model.sample_from_components_no_shuffle(n_add, noise=noise, total=total,
                                                                            max_per_component=mx, row_offset=row_offset,
                                                                            **({} if own is None else {"component_range": own}))
        new_target_grads, new_target_lnpdfs = self.get_target_grads(new_samples)
        return new_samples, new_target_lnpdfs, new_target_grads, mapping

    def select_samples(self, noise=None):
        """sample_selector.py:204-219 -> (samples, mapping, bg, target_lnpdfs, target_grads)  (quirk 2)."""
        num_samples_to_reuse = self.reused_samples_per_component * self.model.num_components
        oldsamples_pdf, samples, _, _, _ = self.sample_db.get_newest_samples(num_samples_to_reuse)
        num_reused_samples = samples.shape[0]
        new_samples, new_target_lnpdfs, new_target_grads, mapping = self.sample_where_needed(samples, oldsamples_pdf,
                                                                                             noise=noise)
        chols = self.model.chol_cov if self.model.shard is None else self.model.chol_cov_handle    # stored, not read
        self.sample_db.add_samples(new_samples, self.model.means, chols, new_target_lnpdfs,
                                   new_target_grads, mapping, prepared=self._prepared())
        num_new_samples = new_samples.shape[0]
        oldsamples_pdf, samples, mapping, target_lnpdfs, target_grads = self.sample_db.get_newest_samples(
            num_reused_samples + num_new_samples)
        return samples, mapping, oldsamples_pdf, target_lnpdfs, target_grads


    def select_samples_deferred(self, noise=None):
        """The no-reuse iteration WITHOUT touching the sample database -> ((samples, mapping, bg, target_lnpdfs,
        target_grads), payload).  With ratio_reused_samples_to_desired = 0 `select_samples` reads back exactly the samples
        it has just stored, and their background density is the mixture they were drawn from with count-proportional
        weights (sample_db.py:217-227); computing that directly keeps every shape and address fixed, which is what a
        captured CUDA graph of the iteration needs (optimization/graphed.py).  `payload` is what `add_samples` has to
        store afterwards (the parameters the samples were drawn from are copied: the update overwrites them).
        `mapping` holds component indices of the current model; the database-global offset of sample_db.py:115 cancels
        in the estimators' relative mapping (ng_estimator.py:244)."""
        if self.reused_samples_per_component != 0:
            raise NotImplementedError("select_samples_deferred needs ratio_reused_samples_to_desired = 0")
        m, db = self.model, self.sample_db
        dev = m.device
        empty = torch.zeros((0, m.num_dimensions), device=dev)
        new_samples, lnpdfs, grads, mapping = self.sample_where_needed(empty, torch.zeros(0, device=dev), noise=noise)
        K = m.num_components
        if db.count_override is not None:
            count = db.count_override.to(torch.float32)
        else:
            count = torch.full((K,), float(max(1, self.desired_samples_per_component)), device=dev)
        weight = count / torch.sum(count)
        if m.diagonal_covs:
            chols = m.chol_cov
            bg, _ = db.evaluate_background(weight, m.means, chols, None, new_samples.contiguous())
            payload = (new_samples, m.means.clone(), chols.clone(), lnpdfs, grads, mapping, None) if db.keep_samples else None
        else:
            linv, _, cst = m.prepared(need_prec=False)
            chols = m.chol_cov if m.shard is None else m.chol_cov_handle
            bg, _ = db.evaluate_background(weight, m.means, chols, linv, new_samples.contiguous(), cst)
            payload = ((new_samples, m.means.clone(), chols.clone(), lnpdfs, grads, mapping,
                        (linv.clone(), None, cst.clone())) if db.keep_samples else None)
        return (new_samples, mapping, bg, lnpdfs, grads), payload


class LinSampleSelector(SampleSelector):
    def __init__(self, target_distribution, model, sample_db, desired_samples_per_component, ratio_reused_samples_to_desired):
        super().__init__(target_distribution, model, sample_db)
        self.desired_samples_per_component = int(desired_samples_per_component)
        self.reused_samples_per_component = int(math.floor(ratio_reused_samples_to_desired * desired_samples_per_component))

    def get_effective_samples(self, model_densities, oldsamples_pdf):
        """sample_selector.py:258-278 (one row: the mixture density)."""
        return ops.importance_weights(model_densities.reshape(1, -1).contiguous(), oldsamples_pdf, want_ess=True)["ess"]

    def sample_where_needed(self, noise=None, uniforms=None):
        """sample_selector.py:280-325 -> (new_samples, mapping, num_reused_samples)."""
        num_samples_to_reuse = self.reused_samples_per_component * self.model.num_components
        oldsamples_pdf, old_samples, _, _, _ = self.sample_db.get_newest_samples(num_samples_to_reuse)
        num_reused_samples = old_samples.shape[0]
        if old_samples.shape[0] == 0:
            n_eff = 0
        else:
            model_logpdfs = self.model.log_density(old_samples)
            n_eff = int(torch.floor(self.get_effective_samples(model_logpdfs, oldsamples_pdf))[0].item())
        n_add = max(1, self.desired_samples_per_component - n_eff)
        new_samples, mapping = self.model.sample(n_add, uniforms=uniforms, noise=noise)
        return new_samples, mapping, num_reused_samples

    def select_samples_deferred(self, noise=None):
        """The no-reuse iteration of the mixture-based selector WITHOUT touching the sample database (see
        VipsSampleSelector.select_samples_deferred): n_eff = 0 without old samples, so n_add = desired
        (sample_selector.py:300-312); the component of every draw comes from the device generator, the per-component
        counts stay on the device (the sampling kernel is launched for the worst case of one component drawing everything);
        the background density weights every component by its share of the draws (sample_db.py:217-227; a component
        that drew nothing has weight 0)."""
        if self.reused_samples_per_component != 0:
            raise NotImplementedError("select_samples_deferred needs ratio_reused_samples_to_desired = 0")
        m, db = self.model, self.sample_db
        n_add = max(1, self.desired_samples_per_component)
        K = m.num_components
        comps = m.sample_categorical(n_add)
        counts = torch.zeros(K, device=m.device, dtype=torch.int32)
        counts.scatter_add_(0, comps.long(), torch.ones_like(comps))
        new_samples, _ = m.sample_from_components_no_shuffle(counts, noise=noise, total=n_add, max_per_component=n_add)
        grads, lnpdfs = self.get_target_grads(new_samples)
        # count / sum(count) with a DEVICE divisor, exactly like get_newest_samples (sample_db.py:226): dividing by a Python
        # scalar is a multiplication by its rounded reciprocal in torch, one ulp off for some counts -- MORE amplifies that
        # to 1e-4 in the updated means (found by the eager-vs-graph comparison, profiles/debug/deferred_vs_eager_probe.py)
        cf = counts.to(torch.float32)
        weight = cf / torch.sum(cf)
        if m.diagonal_covs:
            chols = m.chol_cov
            bg, _ = db.evaluate_background(weight, m.means, chols, None, new_samples.contiguous())
            payload = (new_samples, m.means.clone(), chols.clone(), lnpdfs, grads, comps, None) if db.keep_samples else None
        else:
            linv, _, cst = m.prepared(need_prec=False)
            chols = m.chol_cov
            bg, _ = db.evaluate_background(weight, m.means, chols, linv, new_samples.contiguous(), cst)
            payload = ((new_samples, m.means.clone(), chols.clone(), lnpdfs, grads, comps,
                        (linv.clone(), None, cst.clone())) if db.keep_samples else None)
        return (new_samples, comps, bg, lnpdfs, grads), payload

    def select_samples(self, noise=None, uniforms=None):
        """sample_selector.py:327-339.  `noise` / `uniforms` inject the draws of GMM.sample (parity tests)."""
        new_samples, mapping, num_reused_samples = self.sample_where_needed(noise, uniforms)
        new_target_grads, new_target_lnpdfs = self.get_target_grads(new_samples)
        self.sample_db.add_samples(new_samples, self.model.means, self.model.chol_cov, new_target_lnpdfs,
                                   new_target_grads, mapping, prepared=self._prepared())
        samples_this_iter = num_reused_samples + new_samples.shape[0]
        oldsamples_pdf, samples, mapping, target_lnpdfs, target_grads = self.sample_db.get_newest_samples(samples_this_iter)
        return samples, mapping, oldsamples_pdf, target_lnpdfs, target_grads
