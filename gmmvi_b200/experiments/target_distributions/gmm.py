"""GMM target (mirror of experiments/target_distributions/gmm.py:12-145).  The target is itself a full-covariance
mixture, so its log-density and gradient run on the same device kernels as the model."""
from __future__ import annotations

import numpy as np
import torch

from ...models.full_cov_gmm import FullCovGMM
from .lnpdf import LNPDF


class GMM_LNPDF(LNPDF):
    def __init__(self, target_weights, target_means, target_covs, device="cuda"):
        # the analytic gradient replaces the reference's GradientTape (sample_selector.py:69-78)
        super().__init__(use_log_density_and_grad=True, safe_for_tf_graph=True)
        self.target_weights = torch.as_tensor(np.asarray(target_weights), dtype=torch.float32)
        self.target_means = torch.as_tensor(np.asarray(target_means), dtype=torch.float32)
        self.target_covs = torch.as_tensor(np.asarray(target_covs), dtype=torch.float32)
        self.gmm = FullCovGMM(self.target_weights, self.target_means, self.target_covs, device=device)

    @classmethod
    def from_cholesky(cls, target_weights, target_means, target_chols, device="cuda"):
        """Same target, given the Cholesky factors of the component covariances."""
        self = cls.__new__(cls)
        LNPDF.__init__(self, use_log_density_and_grad=True, safe_for_tf_graph=True)
        self.target_weights = torch.as_tensor(np.asarray(target_weights), dtype=torch.float32)
        self.target_means = torch.as_tensor(np.asarray(target_means), dtype=torch.float32)
        self.gmm = FullCovGMM.from_cholesky(self.target_weights, self.target_means, target_chols, device=device)
        self.target_covs = None
        return self

    def log_density(self, x):
        return self.gmm.log_density(x.to(torch.float32).contiguous())

    def log_density_and_grad(self, x):
        lq, grad, _ = self.gmm.log_density_and_grad(x.to(torch.float32).contiguous())
        return lq, grad

    def marginal_log_density(self, x, dim):
        return self.gmm.marginal_log_density(x.to(torch.float32), dim)

    def get_num_dimensions(self):
        return int(self.target_means.shape[1])

    def can_sample(self):
        return True

    def sample(self, n):
        return self.gmm.sample(n)[0]

    def expensive_metrics(self, model, samples) -> dict:
        """gmm.py:69-121 without the matplotlib figure: number of target modes that have a model mean nearby."""
        tm = self.gmm.means
        d = torch.cdist(tm, model.means).min(dim=1).values
        thr = float(torch.linalg.norm(6.0 * torch.ones(model.num_dimensions)))
        return {"num_detected_modes": int((d < thr).sum().item())}


def make_target(num_dimensions, device="cuda"):
    """experiments/target_distributions/gmm.py:123-145 (NumPy global RNG, like the reference)."""
    num_true_components = 10
    weights = np.ones(num_true_components) / num_true_components
    means = np.empty((num_true_components, num_dimensions))
    covs = np.empty((num_true_components, num_dimensions, num_dimensions))
    for i in range(num_true_components):
        means[i] = 100 * (np.random.random(num_dimensions) - 0.5)
        a = 0.1 * np.random.normal(0, num_dimensions, (num_dimensions * num_dimensions)).reshape(
            (num_dimensions, num_dimensions))
        covs[i] = a.transpose().dot(a) + np.eye(num_dimensions)
    return GMM_LNPDF(weights, means, covs, device=device)


def make_target_with_scale(num_dimensions, num_components, scale, device="cuda"):
    weights = np.ones(num_components) / num_components
    means = np.empty((num_components, num_dimensions))
    covs = np.empty((num_components, num_dimensions, num_dimensions))
    for i in range(num_components):
        means[i] = 100 * (np.random.random(num_dimensions) - 0.5)
        a = np.random.normal(0, np.sqrt(scale), (num_dimensions * num_dimensions)).reshape((num_dimensions, num_dimensions))
        covs[i] = a.transpose().dot(a) + np.eye(num_dimensions)
    return GMM_LNPDF(weights, means, covs, device=device)
