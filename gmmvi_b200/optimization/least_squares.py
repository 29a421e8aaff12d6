"""Weighted ridge regression on quadratic features (mirror of optimization/least_squares.py), used by MORE."""
from __future__ import annotations

import torch

from .. import ops


class QuadFunc:
    """least_squares.py:93-191.  `fit_quadratic_batched` fits all K components at once on device."""

    def __init__(self, dim: int):
        self.dim = dim
        self.num_quad_features = dim * (dim + 1) // 2
        self.num_features = self.num_quad_features + dim + 1

    def fit_quadratic_batched(self, regularizers, samples, rewards, weights, means, linv):
        """least_squares.py:126-191 for all components: -> (quad_term[K,D,D], lin_term[K,D], ok[K])."""
        return ops.more_fit(regularizers, samples, rewards, weights, means, linv)
