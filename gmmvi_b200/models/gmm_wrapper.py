"""GmmWrapper: per-component learner metadata around a GMM (mirror of models/gmm_wrapper.py:4-182)."""
from __future__ import annotations

import torch

from .gmm import GMM

FLT_MIN = float(torch.finfo(torch.float32).min)     # tf.float32.min


class GmmWrapper:
    @staticmethod
    def build_from_config(model: GMM, config: dict):
        """models/gmm_wrapper.py:34-58."""
        nca = config["num_component_adapter_config"]
        max_reward_history_length = 2 * max(2, nca["del_iters"]) if "del_iters" in nca else 2
        initial_regularizer = config["ng_estimator_config"].get("initial_l2_regularizer", 1e-12)
        initial_stepsize = config["component_stepsize_adapter_config"]["initial_stepsize"]
        return GmmWrapper(model, initial_stepsize, initial_regularizer, max_reward_history_length)

    def __init__(self, model: GMM, initial_stepsize: float, initial_regularizer: float, max_reward_history_length: int):
        """models/gmm_wrapper.py:60-81."""
        self.model = model
        self.initial_regularizer = initial_regularizer
        self.initial_last_eta = -1
        self.initial_stepsize = initial_stepsize
        self.max_reward_history_length = max_reward_history_length
        K, dev = model.num_components, model.device
        f = lambda v: torch.full((K,), float(v), device=dev, dtype=torch.float32)
        self.l2_regularizers = f(initial_regularizer)
        self.last_log_etas = f(self.initial_last_eta)
        self.num_received_updates = f(0.0)
        self.stepsizes = f(initial_stepsize)
        self.reward_history = torch.full((K, max_reward_history_length), FLT_MIN, device=dev)
        self.weight_history = torch.full((K, max_reward_history_length), FLT_MIN, device=dev)
        self.unique_component_ids = torch.arange(K, device=dev, dtype=torch.int32)
        self.max_component_id = K - 1
        self.adding_thresholds = f(-1.0)
        self.initial_entropies = model.component_entropies()

    def __getattr__(self, name):
        """Forward everything else to the wrapped model (models/gmm_wrapper.py:83-88)."""
        if name == "model":
            raise AttributeError(name)
        return getattr(self.model, name)

    def add_component(self, initial_weight, initial_mean, initial_cov, adding_threshold, initial_entropy):
        """models/gmm_wrapper.py:90-127."""
        dev = self.model.device
        self.model.add_component(initial_weight, initial_mean, initial_cov)
        self.max_component_id += 1
        one = lambda v: torch.full((1,), float(v), device=dev, dtype=torch.float32)
        H = self.max_reward_history_length
        self.unique_component_ids = torch.cat((self.unique_component_ids,
                                               torch.tensor([self.max_component_id], device=dev, dtype=torch.int32)))
        self.l2_regularizers = torch.cat((self.l2_regularizers, one(self.initial_regularizer)))
        self.last_log_etas = torch.cat((self.last_log_etas, one(self.initial_last_eta)))
        self.num_received_updates = torch.cat((self.num_received_updates, one(0.0)))
        self.stepsizes = torch.cat((self.stepsizes, one(self.initial_stepsize)))
        self.reward_history = torch.cat((self.reward_history, torch.full((1, H), FLT_MIN, device=dev)), 0)
        self.weight_history = torch.cat((self.weight_history, torch.full((1, H), float(initial_weight), device=dev)), 0)
        self.adding_thresholds = torch.cat((self.adding_thresholds,
                                            torch.as_tensor(adding_threshold, dtype=torch.float32, device=dev).reshape(1)))
        self.initial_entropies = torch.cat((self.initial_entropies,
                                            torch.as_tensor(initial_entropy, dtype=torch.float32, device=dev).reshape(1)))

    def remove_component(self, idx: int):
        """models/gmm_wrapper.py:129-148."""
        idx = int(idx)
        self.model.remove_component(idx)
        K = self.l2_regularizers.shape[0]
        sel = torch.tensor([i for i in range(K) if i != idx], device=self.model.device, dtype=torch.long)
        for name in ("unique_component_ids", "l2_regularizers", "last_log_etas", "num_received_updates", "stepsizes",
                     "reward_history", "weight_history", "adding_thresholds", "initial_entropies"):
            setattr(self, name, getattr(self, name)[sel].contiguous())

    def store_rewards(self, rewards: torch.Tensor):
        """models/gmm_wrapper.py:150-158."""
        self.reward_history = torch.cat((self.reward_history[:, 1:], rewards.unsqueeze(1)), dim=1)

    def update_stepsizes(self, new_stepsizes: torch.Tensor):
        self.stepsizes = new_stepsizes.contiguous()

    def replace_weights(self, new_log_weights: torch.Tensor):
        """models/gmm_wrapper.py:170-182."""
        self.model.replace_weights(new_log_weights)
        self.weight_history = torch.cat((self.weight_history[:, 1:], self.model.weights.unsqueeze(1)), dim=1)
