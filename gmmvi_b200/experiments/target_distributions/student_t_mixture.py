"""Mixture of multivariate Student-t targets (mirror of experiments/target_distributions/student_t_mixture.py).
log t_j(x) = lgamma((nu+D)/2) - lgamma(nu/2) - D/2 log(nu pi) - sum log diag L_j - (nu+D)/2 log(1 + m_j/nu),
m_j = |L_j^-1 (x - mu_j)|^2 taken from the Gaussian log-density kernel."""
from __future__ import annotations

from math import lgamma, log, pi

import numpy as np
import torch

from ... import ops
from ...models.full_cov_gmm import FullCovGMM
from .lnpdf import LNPDF


class StudentTMixture_LNPDF(LNPDF):
    def __init__(self, target_weights, target_means, target_covs, alpha=2, device="cuda"):
        super().__init__(use_log_density_and_grad=True)
        self.alpha = alpha
        self.target_weights = torch.as_tensor(np.asarray(target_weights), dtype=torch.float32)
        self.target_means = torch.as_tensor(np.asarray(target_means), dtype=torch.float32)
        self.target_covs = torch.as_tensor(np.asarray(target_covs), dtype=torch.float32)
        self._g = FullCovGMM(self.target_weights, self.target_means, self.target_covs, device=device)
        D = self.get_num_dimensions()
        self._norm = lgamma((alpha + D) / 2) - lgamma(alpha / 2) - 0.5 * D * log(alpha * pi)

    def _component_terms(self, x):
        g = self._g
        x = x.to(torch.float32).contiguous()
        _, _, cst = g.prepared(need_prec=False)
        lq = g.component_log_densities(x)                       # cst_j - m_j / 2
        D, nu = g.num_dimensions, float(self.alpha)
        logdet_part = cst + 0.5 * D * log(2 * pi)               # - sum log diag L_j
        maha = torch.clamp(-2.0 * (lq - cst.unsqueeze(1)), min=0.0)
        lt = self._norm + logdet_part.unsqueeze(1) - 0.5 * (nu + D) * torch.log1p(maha / nu)
        return x, lt.contiguous(), maha

    def log_density(self, x):
        _, lt, _ = self._component_terms(x)
        return ops.mixture_lse(lt, self._g.log_weights)

    def log_density_and_grad(self, x):
        g = self._g
        x, lt, maha = self._component_terms(x)
        logp = ops.mixture_lse(lt, g.log_weights)
        D, nu = g.num_dimensions, float(self.alpha)
        # grad = - sum_j r_j (nu + D)/(nu + m_j) P_j (x - mu_j): responsibilities folded into the "log density" operand
        lr = lt + g.log_weights.unsqueeze(1) - logp.unsqueeze(0) + torch.log((nu + D) / (nu + maha))
        zero_w = torch.zeros_like(g.log_weights)
        _, prec, _ = g.prepared()
        grad = ops.mixture_grad_full(x, g.means, prec, lr.contiguous(), zero_w, torch.zeros_like(logp))
        return logp, grad

    def get_num_dimensions(self):
        return int(self.target_means.shape[1])


def make_target(num_dimensions, harder_setting, use_matlab_target=False, device="cuda"):
    """student_t_mixture.py:138-194 (the MATLAB data files are not shipped with the reference)."""
    if use_matlab_target:
        raise ValueError("Matlab data is not available for the mixture of Student-T experiment")
    s, num_components = (25, 20) if harder_setting else (20, 10)
    weights = np.ones(num_components) / num_components
    means = np.empty((num_components, num_dimensions))
    covs = np.empty((num_components, num_dimensions, num_dimensions))
    for i in range(num_components):
        means[i] = torch.rand(num_dimensions).numpy() * (2 * s) - s
        a = 0.1 * num_dimensions * np.random.normal(0, 1, (num_dimensions * num_dimensions)).reshape(
            (num_dimensions, num_dimensions))
        covs[i] = np.linalg.inv(a.transpose().dot(a) + np.eye(num_dimensions))
    return StudentTMixture_LNPDF(weights, means, covs, device=device)
