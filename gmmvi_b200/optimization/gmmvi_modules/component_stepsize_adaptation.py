"""Per-component stepsize schedules (mirror of optimization/gmmvi_modules/component_stepsize_adaptation.py).
O(K) element-wise glue on device vectors; no kernel."""
from __future__ import annotations

import torch


class ComponentStepsizeAdaptation:
    def __init__(self, gmm_wrapper, initial_stepsize: float):
        self.gmm_wrapper = gmm_wrapper
        self.initial_stepsize = initial_stepsize
        if not torch.all(gmm_wrapper.stepsizes == torch.full_like(gmm_wrapper.stepsizes, float(initial_stepsize))):
            raise AssertionError("gmm_wrapper.stepsizes != initial_stepsize")     # :28

    @staticmethod
    def build_from_config(config, gmm_wrapper):
        """component_stepsize_adaptation.py:30-52."""
        t = config["component_stepsize_adapter_type"]
        cfg = config["component_stepsize_adapter_config"]
        if t == "improvement-based":
            return ImprovementBasedComponentStepsizeAdaptation(gmm_wrapper, **cfg)
        elif t == "decaying":
            return DecayingComponentStepsizeAdaptation(gmm_wrapper, **cfg)
        elif t == "fixed":
            return FixedComponentStepsizeAdaptation(gmm_wrapper, **cfg)
        raise ValueError(f"config['component_stepsize_adapter_type'] is '{t}' which is an unknown type")

    def update_stepsize(self, current_stepsizes):
        raise NotImplementedError


class FixedComponentStepsizeAdaptation(ComponentStepsizeAdaptation):
    def update_stepsize(self, current_stepsizes):
        return current_stepsizes


class DecayingComponentStepsizeAdaptation(ComponentStepsizeAdaptation):
    def __init__(self, gmm_wrapper, annealing_exponent: float, initial_stepsize: float):
        super().__init__(gmm_wrapper, initial_stepsize)
        self.annealing_exponent = annealing_exponent

    def update_stepsize(self, current_stepsizes):
        """:116-130."""
        n = self.gmm_wrapper.num_received_updates
        return self.initial_stepsize / (1 + torch.pow(n, float(self.annealing_exponent)))


class ImprovementBasedComponentStepsizeAdaptation(ComponentStepsizeAdaptation):
    def __init__(self, gmm_wrapper, initial_stepsize, min_stepsize, max_stepsize, stepsize_inc_factor, stepsize_dec_factor):
        super().__init__(gmm_wrapper, initial_stepsize)
        self.min_stepsize, self.max_stepsize = min_stepsize, max_stepsize
        self.stepsize_inc_factor, self.stepsize_dec_factor = stepsize_inc_factor, stepsize_dec_factor

    def update_stepsize(self, current_stepsizes):
        """:165-188 (decrease when history[-2] >= history[-1], quirk 14)."""
        h = self.gmm_wrapper.reward_history
        worse = h[:, -2] >= h[:, -1]
        dec = torch.clamp(self.stepsize_dec_factor * current_stepsizes, min=float(self.min_stepsize))
        inc = torch.clamp(self.stepsize_inc_factor * current_stepsizes, max=float(self.max_stepsize))
        return torch.where(worse, dec, inc)
