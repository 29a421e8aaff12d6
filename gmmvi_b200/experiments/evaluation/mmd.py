"""Maximum Mean Discrepancy between a model sample and a ground-truth sample (mirror of
experiments/evaluation/mmd.py:4-78).  Same class surface: MMD(groundtruth, alpha), compute_sigma, compute_ustat,
kernel_mix, set_alpha, compute_MMD.  The O(n^2 D) Gaussian-kernel sums run in `gvi_gauss_kernel_sum_f32`
(gmmvi_b200/csrc/elementwise.cu); the median trick is O(P^2 D log P) torch glue that runs once per object."""
from __future__ import annotations

import torch

from ... import ops


class MMD:
    def __init__(self, groundtruth, alpha, device="cuda"):
        """mmd.py:20-24."""
        self.groundtruth = torch.as_tensor(groundtruth, dtype=torch.float32).to(device).contiguous()
        self.num_groundtruth = int(self.groundtruth.shape[0])
        self.sigma = self.compute_sigma()
        self.set_alpha(alpha)

    def compute_sigma(self, max_points_for_median=1000):
        """mmd.py:26-36: diagonal bandwidth = per-dimension median of the squared differences of all pairs i <= j
        (the i == j zeros included) of the first `max_points_for_median` ground-truth points.  The median is
        tfp.stats.percentile(., 50) with its default 'nearest' interpolation: element round_half_even((d - 1) / 2) of
        the DESCENDING sort."""
        P = min(int(max_points_for_median), self.num_groundtruth)
        G = self.groundtruth[:P]
        iu = torch.triu_indices(P, P, device=G.device)
        d = int(iu.shape[1])
        k_desc = int(round((d - 1) * 0.5))                      # Python's round is round-half-even, like tf.round
        k_asc = d - 1 - k_desc
        D = G.shape[1]
        med = torch.empty(D, device=G.device, dtype=torch.float32)
        step = max(1, (1 << 27) // max(d, 1))                      # dimensions per chunk: <= 512 MB of differences
        for d0 in range(0, D, step):
            cols = G[:, d0:d0 + step]
            diff = (cols[iu[0]] - cols[iu[1]]) ** 2
            med[d0:d0 + step] = torch.kthvalue(diff, k_asc + 1, dim=0).values
        return torch.diag(med)

    def compute_ustat(self, sample, alpha):
        """mmd.py:38-46: sum_{i,j} exp(-(s_i - s_j)^T (alpha sigma)^-1 (s_i - s_j))."""
        sample = torch.as_tensor(sample, dtype=torch.float32, device=self.groundtruth.device).contiguous()
        return ops.gauss_kernel_sum(sample, sample, self._bandwidth(alpha))

    def kernel_mix(self, sample, alpha):
        """mmd.py:48-56: the same sum over (ground-truth point, sample point) pairs."""
        sample = torch.as_tensor(sample, dtype=torch.float32, device=self.groundtruth.device).contiguous()
        return ops.gauss_kernel_sum(self.groundtruth, sample, self._bandwidth(alpha))

    def _bandwidth(self, alpha):
        # tf.linalg.inv(alpha * sigma) of the diagonal matrix (a zero median gives inf like the reference's inverse)
        return (1.0 / (float(alpha) * torch.diagonal(self.sigma))).contiguous()

    def set_alpha(self, alpha):
        """mmd.py:58-60."""
        self._alpha = alpha
        self.ustat1 = self.compute_ustat(self.groundtruth, alpha)

    def compute_MMD(self, model_sample):
        """mmd.py:62-78 -> 0-d float32 tensor."""
        num_1 = self.num_groundtruth
        num_2 = int(model_sample.shape[0])
        mmd = self.ustat1 / (num_1 ** 2) \
            + self.compute_ustat(model_sample, self._alpha) / (num_2 ** 2) \
            - 2 * self.kernel_mix(model_sample, self._alpha) / (num_1 * num_2)
        return mmd
