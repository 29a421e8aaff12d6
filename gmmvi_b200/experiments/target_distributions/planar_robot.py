"""Planar-robot target (mirror of experiments/target_distributions/planar_robot.py:13-138): D = num_links joint
angles, zero-mean Gaussian prior, max over goal Gaussians on the end-effector position.  The reference differentiates
`log_density` with a GradientTape (use_log_density_and_grad=False, sample_selector.py:73-77); here density and gradient
come from one fused kernel (`gvi_planar_robot_f32`, analytic gradient through the arg-max goal), so the target reports
use_log_density_and_grad=True.  `likelihood` / `forward_kinematics` keep the reference's signatures (torch ops; they are
only used by the metrics)."""
from __future__ import annotations

from math import log, pi

import numpy as np
import torch

from ... import ops
from .lnpdf import LNPDF


class PlanarRobot(LNPDF):
    def __init__(self, num_links, num_goals, prior_std=2e-1, likelihood_std=1e-2, device="cuda"):
        super().__init__(use_log_density_and_grad=True)
        self._num_dimensions = num_links
        prior_stds = prior_std * np.ones(num_links)
        prior_stds[0] = 1.0
        self.prior_stds = torch.tensor(prior_stds, dtype=torch.float32, device=device)
        self.link_lengths = np.ones(num_links)
        self._num_goals = num_goals
        if num_goals == 1:
            goals = [[7.0, 0.0]]
        elif num_goals == 4:
            goals = [[7.0, 0.0], [-7.0, 0.0], [0.0, 7.0], [0.0, -7.0]]
        else:
            raise ValueError
        self.goals = torch.tensor(goals, dtype=torch.float32, device=device)
        self.likelihood_std = likelihood_std

    def likelihood(self, pos):
        d = pos.unsqueeze(0) - self.goals.unsqueeze(1)                     # [G, N, 2]
        lp = -0.5 * torch.sum((d / self.likelihood_std) ** 2, dim=2) - 2 * log(self.likelihood_std) - log(2 * pi)
        return torch.max(lp, dim=0).values

    def get_num_dimensions(self):
        return self._num_dimensions

    def forward_kinematics(self, theta):
        cs = torch.cumsum(theta, dim=1)
        ll = torch.as_tensor(self.link_lengths, dtype=theta.dtype, device=theta.device)
        x = torch.sum(ll * torch.cos(cs), dim=1)
        y = torch.sum(ll * torch.sin(cs), dim=1)
        return torch.stack((x, y), dim=1)

    def log_density(self, theta):
        """planar_robot.py:65-66."""
        return ops.planar_robot(theta.to(torch.float32), self.prior_stds, self.goals, self.likelihood_std, False)[0]

    def log_density_and_grad(self, theta):
        return ops.planar_robot(theta.to(torch.float32), self.prior_stds, self.goals, self.likelihood_std, True)


def make_single_goal(device="cuda"):
    return PlanarRobot(10, 1, device=device)


def make_four_goal(device="cuda"):
    return PlanarRobot(10, 4, device=device)
