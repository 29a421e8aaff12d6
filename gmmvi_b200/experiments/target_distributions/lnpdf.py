"""Target-distribution interface (mirror of experiments/target_distributions/lnpdf.py:6-127).
`x` is a CUDA torch tensor [N, D]; densities are fp32 tensors [N]."""
from __future__ import annotations


class LNPDF:
    def __init__(self, use_log_density_and_grad: bool = False, safe_for_tf_graph: bool = True):
        self._use_log_density_and_grad = use_log_density_and_grad
        self._safe_for_tf_graph = safe_for_tf_graph

    def log_density(self, x):
        raise NotImplementedError

    def log_density_and_grad(self, x):
        raise NotImplementedError

    def get_num_dimensions(self) -> int:
        raise NotImplementedError

    def expensive_metrics(self, model, samples) -> dict:
        return dict()

    def can_sample(self) -> bool:
        return False

    @property
    def use_log_density_and_grad(self) -> bool:
        return self._use_log_density_and_grad

    @property
    def safe_for_tf_graph(self) -> bool:
        return self._safe_for_tf_graph

    def sample(self, n: int):
        raise NotImplementedError
