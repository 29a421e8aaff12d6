"""Weight stepsize schedules (mirror of optimization/gmmvi_modules/weight_stepsize_adaptation.py).
The stepsize lives in a device scalar so that no schedule forces a host synchronisation."""
from __future__ import annotations

import torch

FLT_MIN = float(torch.finfo(torch.float32).min)


class WeightStepsizeAdaptation:
    def __init__(self, initial_stepsize, device="cuda"):
        self.stepsize = torch.tensor([float(initial_stepsize)], device=device, dtype=torch.float32)

    @staticmethod
    def build_from_config(config, gmm_wrapper):
        """weight_stepsize_adaptation.py:26-48 (type string is 'improvement_based', with an underscore)."""
        t = config["weight_stepsize_adapter_type"]
        cfg = config["weight_stepsize_adapter_config"]
        dev = gmm_wrapper.device
        if t == "fixed":
            return FixedWeightStepsizeAdaptation(device=dev, **cfg)
        elif t == "decaying":
            return DecayingWeightStepsizeAdaptation(device=dev, **cfg)
        elif t == "improvement_based":
            return ImprovementBasedWeightStepsizeAdaptation(gmm_wrapper, **cfg)
        raise ValueError(f"config['weight_stepsize_adapter_type'] is '{t}' which is an unknown type")

    def _update_stepsize(self):
        pass

    def update_stepsize(self):
        self._update_stepsize()
        return self.stepsize


class FixedWeightStepsizeAdaptation(WeightStepsizeAdaptation):
    pass


class DecayingWeightStepsizeAdaptation(WeightStepsizeAdaptation):
    def __init__(self, initial_stepsize, annealing_exponent, device="cuda"):
        super().__init__(initial_stepsize, device)
        self.initial_stepsize = float(initial_stepsize)
        self.annealing_exponent = float(annealing_exponent)
        self.num_weight_updates = torch.zeros(1, device=device, dtype=torch.float32)     # device counter (graph-safe)

    def _update_stepsize(self):
        """:96-105."""
        self.stepsize = self.initial_stepsize / (1.0 + torch.pow(self.num_weight_updates, self.annealing_exponent))
        self.num_weight_updates = self.num_weight_updates + 1.0


class ImprovementBasedWeightStepsizeAdaptation(WeightStepsizeAdaptation):
    def __init__(self, model, initial_stepsize, min_stepsize, max_stepsize, stepsize_inc_factor, stepsize_dec_factor):
        super().__init__(initial_stepsize, model.device)
        self.model = model
        self.min_stepsize, self.max_stepsize = float(min_stepsize), float(max_stepsize)
        self.stepsize_inc_factor, self.stepsize_dec_factor = float(stepsize_inc_factor), float(stepsize_dec_factor)
        self.elbo_history = torch.tensor([FLT_MIN], device=model.device)

    def _update_stepsize(self):
        """:141-156."""
        m = self.model
        w = m.weights
        elbo = torch.sum(w * m.reward_history[:, -1]) - torch.sum(w * m.log_weights)
        self.elbo_history = torch.cat((self.elbo_history[-1:], elbo.reshape(1)))
        better = self.elbo_history[-1] > self.elbo_history[-2]
        inc = torch.clamp(self.stepsize_inc_factor * self.stepsize, max=self.max_stepsize)
        dec = torch.clamp(self.stepsize_dec_factor * self.stepsize, min=self.min_stepsize)
        self.stepsize = torch.where(better, inc, dec)
