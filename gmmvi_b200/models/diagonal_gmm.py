"""Diagonal-covariance GMM (mirror of models/diagonal_gmm.py:6-59); chol_cov holds std-devs [K, D]."""
from __future__ import annotations

from math import log, pi

import torch

from .. import ops
from .gmm import GMM, _as_param


class DiagonalGMM(GMM):
    def __init__(self, weights, means, covs, device="cuda"):
        means = _as_param(means, device)
        chol = torch.sqrt(_as_param(covs, device))
        log_weights = torch.log(_as_param(weights, device))
        super().__init__(log_weights, means, chol)
        self.diagonal_covs = True

    @staticmethod
    def diagonal_gaussian_log_pdf(dim: int, mean: torch.Tensor, chol: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """models/diagonal_gmm.py:31-34 (single component; routed through the batched kernel)."""
        return ops.logdens_diag(x, mean.reshape(1, -1).contiguous(), chol.reshape(1, -1).contiguous())[0]

    @property
    def covs(self) -> torch.Tensor:
        return torch.square(self.chol_cov)

    def gaussian_entropy(self, chol: torch.Tensor) -> torch.Tensor:
        return 0.5 * self.num_dimensions * (log(2 * pi) + 1) + torch.sum(torch.log(chol))

    def component_log_densities(self, samples: torch.Tensor) -> torch.Tensor:
        """models/diagonal_gmm.py:47-53 -> [K, N]."""
        return ops.logdens_diag(samples, self.means, self.chol_cov)

    def _mixture_grad(self, samples, lq, logq, logw=None, index=None):
        means, stds = self.means, self.chol_cov
        if index is not None:
            means, stds = means[index:index + 1].contiguous(), stds[index:index + 1].contiguous()
        return ops.mixture_grad_diag(samples, means, stds, lq, self.log_weights if logw is None else logw, logq)

    def add_component(self, initial_weight, initial_mean, initial_cov):
        """models/diagonal_gmm.py:55-59."""
        dev = self.device
        self.means = torch.cat((self.means, torch.as_tensor(initial_mean, dtype=torch.float32, device=dev).reshape(1, -1)), 0)
        self.chol_cov = torch.cat((self.chol_cov, torch.sqrt(torch.as_tensor(initial_cov, dtype=torch.float32, device=dev)).reshape(1, -1)), 0).contiguous()
        w = torch.log(torch.as_tensor(initial_weight, dtype=torch.float32, device=dev).reshape(1))
        self.replace_weights(torch.cat((self.log_weights, w), 0))
