"""Full-covariance GMM (mirror of models/full_cov_gmm.py:6-67)."""
from __future__ import annotations

from math import log, pi

import torch

from .. import ops
from .gmm import GMM, _as_param


def _batched_cholesky(covs: torch.Tensor, device) -> torch.Tensor:
    """Factorisation of user-supplied covariances (models/full_cov_gmm.py:23, :67: the constructor and add_component,
    which SAMTRON calls every `add_iters` iterations): `gvi_cholesky_f32`, fp64 arithmetic on the device rounded to fp32,
    NaN factor for a matrix that is not positive definite like tf.linalg.cholesky.  No host round trip."""
    covs = torch.as_tensor(covs, dtype=torch.float32).to(device).contiguous()
    return ops.cholesky(covs)[0]


class FullCovGMM(GMM):
    def __init__(self, weights, means, covs, device="cuda"):
        self.diagonal_covs = False
        means = _as_param(means, device)
        chol = _batched_cholesky(covs, device)
        log_weights = torch.log(_as_param(weights, device))
        super().__init__(log_weights, means, chol)

    @classmethod
    def from_cholesky(cls, weights, means, chols, device="cuda"):
        """Build the mixture directly from lower-triangular Cholesky factors [K,D,D] (no host factorisation)."""
        self = cls.__new__(cls)
        self.diagonal_covs = False
        GMM.__init__(self, torch.log(_as_param(weights, device)), _as_param(means, device), _as_param(chols, device))
        return self

    @property
    def covs(self) -> torch.Tensor:
        """models/full_cov_gmm.py:29-31."""
        return ops.bgemm(self.chol_cov, self.chol_cov, False, True)

    def gaussian_entropy(self, chol: torch.Tensor) -> torch.Tensor:
        """models/full_cov_gmm.py:33-34."""
        return 0.5 * self.num_dimensions * (log(2 * pi) + 1) + torch.sum(torch.log(torch.diagonal(chol)))

    def component_log_density(self, index: int, samples: torch.Tensor) -> torch.Tensor:
        """models/full_cov_gmm.py:41-47."""
        linv, _, cst = self.prepared(need_prec=False)
        i = int(index)
        return ops.logdens_full(samples, self.means[i:i + 1].contiguous(), linv[i:i + 1].contiguous(),
                                cst[i:i + 1].contiguous())[0]

    def component_marginal_log_densities(self, samples: torch.Tensor, dim: int) -> torch.Tensor:
        """models/full_cov_gmm.py:49-54."""
        var = torch.sum(self.chol_cov[:, dim, :] ** 2, dim=1)          # covs[:, dim, dim]
        diffs = samples[:, dim].unsqueeze(0) - self.means[:, dim].unsqueeze(1)
        mahalas = -0.5 * diffs * diffs / var.unsqueeze(1)
        const_parts = -0.5 * torch.log(var) - 0.5 * log(2 * pi)
        return mahalas + const_parts.unsqueeze(1)

    def component_log_densities(self, samples: torch.Tensor) -> torch.Tensor:
        """models/full_cov_gmm.py:56-62 -> [K, N]."""
        linv, _, cst = self.prepared(need_prec=False)
        return ops.logdens_full(samples, self.means, linv, cst)

    def _mixture_grad(self, samples, lq, logq, logw=None, index=None):
        _, prec, _ = self.prepared()
        means = self.means
        if index is not None:
            means, prec = means[index:index + 1].contiguous(), prec[index:index + 1].contiguous()
        return ops.mixture_grad_full(samples, means, prec, lq, self.log_weights if logw is None else logw, logq)

    def add_component(self, initial_weight, initial_mean, initial_cov):
        """models/full_cov_gmm.py:64-67."""
        dev = self.device
        new_chol = _batched_cholesky(torch.as_tensor(initial_cov, dtype=torch.float32).reshape(1, self.num_dimensions, -1), dev)
        self.means = torch.cat((self.means, torch.as_tensor(initial_mean, dtype=torch.float32, device=dev).reshape(1, -1)), 0)
        self.chol_cov = torch.cat((self.chol_cov, new_chol), 0).contiguous()
        w = torch.log(torch.as_tensor(initial_weight, dtype=torch.float32, device=dev).reshape(1))
        self.replace_weights(torch.cat((self.log_weights, w), 0))
