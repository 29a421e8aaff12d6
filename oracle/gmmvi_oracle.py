"""NumPy restatement of the reference's SAMTRON hot path (TEST INFRASTRUCTURE ONLY).

PARITY PIN (see oracle/__init__.py): the reference has no golden vectors and TF cannot be imported here; every function
below follows the cited reference lines op for op and is pinned (a) against the outputs of the reference's own sources
executed over a torch stand-in for the TensorFlow API (tests/golden/reference_*.npz, <= 1e-13 in float64) and (b) against
independent closed forms, both in tests/test_oracle_pins.py.

All paths are relative to /root/reference/src/gmmvi/ .  `dt` selects the arithmetic type:
np.float32 mimics the reference (tf.float32 everywhere), np.float64 is the "truth" used to
measure the fp32 noise floor.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np
from scipy.linalg import solve_triangular

FLT_MAX = float(np.finfo(np.float32).max)       # tf.float32.max
FLT_MIN = -FLT_MAX                               # tf.float32.min (most negative finite)
LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# TF primitives restated
# --------------------------------------------------------------------------------------
def logsumexp(a, axis=None, keepdims=False):
    """tf.reduce_logsumexp: max-shifted, max replaced by 0 when not finite."""
    a = np.asarray(a)
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0).astype(a.dtype)
    with np.errstate(divide="ignore"):
        out = np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True)) + m
    if not keepdims:
        out = np.squeeze(out, axis=axis) if axis is not None else out.reshape(())
    return out.astype(a.dtype)


def cholesky_or_nan(a):
    """tf.linalg.cholesky: a failed factorisation yields NaNs instead of raising (TF >= 2.5)."""
    try:
        if not np.all(np.isfinite(a)):
            raise np.linalg.LinAlgError
        return np.linalg.cholesky(a)
    except np.linalg.LinAlgError:
        return np.full_like(a, np.nan)


def unique_with_counts_first_occurrence(x):
    """tf.unique_with_counts: unique values in order of first occurrence, idx, counts."""
    x = np.asarray(x)
    vals, first, inv, cnt = np.unique(x, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return vals[order], rank[inv].astype(np.int32), cnt[order]


def reduce_weighted_logsumexp_sign(logx, w, axis):
    """tfp.math.reduce_weighted_logsumexp(logx, w, axis, return_sign=True)."""
    m = np.max(logx, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0).astype(logx.dtype)
    s = np.sum(w * np.exp(logx - m), axis=axis, keepdims=True)
    sign = np.sign(s)
    with np.errstate(divide="ignore"):
        lswe = np.log(np.abs(s)) + m
    return np.squeeze(lswe, axis=axis), np.squeeze(sign, axis=axis)


# --------------------------------------------------------------------------------------
# Model state (models/gmm.py:27-34 + models/gmm_wrapper.py:60-81)
# --------------------------------------------------------------------------------------
@dataclass
class OracleGMM:
    log_weights: np.ndarray          # [K]
    means: np.ndarray                # [K, D]
    chol_cov: np.ndarray             # [K, D, D] lower  (diag: [K, D] std-devs)
    diagonal_covs: bool = False
    # GmmWrapper metadata
    initial_stepsize: float = 1.0
    initial_regularizer: float = 1e-12
    max_reward_history_length: int = 2
    l2_regularizers: np.ndarray = field(default=None)
    last_log_etas: np.ndarray = field(default=None)
    num_received_updates: np.ndarray = field(default=None)
    stepsizes: np.ndarray = field(default=None)
    reward_history: np.ndarray = field(default=None)
    weight_history: np.ndarray = field(default=None)

    def __post_init__(self):
        dt = self.means.dtype
        K = self.means.shape[0]
        self.replace_weights(self.log_weights)          # gmm.py:34
        if self.l2_regularizers is None:                # gmm_wrapper.py:68-74
            self.l2_regularizers = np.full(K, self.initial_regularizer, dt)
            self.last_log_etas = np.full(K, -1.0, dt)
            self.num_received_updates = np.zeros(K, dt)
            self.stepsizes = np.full(K, self.initial_stepsize, dt)
            self.reward_history = np.full((K, self.max_reward_history_length), FLT_MIN, dt)
            self.weight_history = np.full((K, self.max_reward_history_length), FLT_MIN, dt)

    @property
    def dt(self):
        return self.means.dtype

    @property
    def num_components(self):
        return self.means.shape[0]

    @property
    def num_dimensions(self):
        return self.means.shape[1]

    @property
    def weights(self):
        return np.exp(self.log_weights)

    def replace_weights(self, new_log_weights):
        """gmm.py:173-181 (+ gmm_wrapper.py:170-182 history shift done by caller)."""
        v = np.asarray(new_log_weights, self.means.dtype)
        self.log_weights = v - logsumexp(v)

    def wrapper_replace_weights(self, new_log_weights):
        """gmm_wrapper.py:170-182."""
        self.replace_weights(new_log_weights)
        self.weight_history = np.concatenate((self.weight_history[:, 1:], self.weights[:, None]), axis=1)

    def store_rewards(self, rewards):
        """gmm_wrapper.py:150-158."""
        self.reward_history = np.concatenate(
            (self.reward_history[:, 1:], np.asarray(rewards, self.dt)[:, None]), axis=1)


def make_full_gmm(weights, means, covs, dt=np.float32, **kw) -> OracleGMM:
    """models/full_cov_gmm.py:19-27."""
    covs = np.asarray(covs, dt)
    chol = np.stack([np.linalg.cholesky(c) for c in covs]).astype(dt)
    return OracleGMM(np.log(np.asarray(weights, dt)), np.asarray(means, dt), chol, False, **kw)


def make_diag_gmm(weights, means, diag_covs, dt=np.float32, **kw) -> OracleGMM:
    """models/diagonal_gmm.py:20-28."""
    return OracleGMM(np.log(np.asarray(weights, dt)), np.asarray(means, dt),
                     np.sqrt(np.asarray(diag_covs, dt)), True, **kw)


# --------------------------------------------------------------------------------------
# A1 / A4: component log densities
# --------------------------------------------------------------------------------------
def full_component_log_densities(X, means, chols):
    """models/full_cov_gmm.py:56-62 -> [K, N]."""
    dt = X.dtype
    K, D = means.shape
    out = np.empty((K, X.shape[0]), dt)
    for k in range(K):
        diffs = X - means[k]                                                   # :57
        sqrts = solve_triangular(chols[k], diffs.T, lower=True, check_finite=False)  # :58
        mahalas = -0.5 * np.sum(sqrts * sqrts, axis=0)                         # :59
        const = -0.5 * np.sum(np.log(np.square(np.diag(chols[k])))) - dt.type(0.5 * D * LOG_2PI)  # :60-61
        out[k] = (mahalas + const).astype(dt)
    return out


def diag_component_log_densities(X, means, stds):
    """models/diagonal_gmm.py:31-34,47-53 -> [K, N]."""
    dt = X.dtype
    K, D = means.shape
    out = np.empty((K, X.shape[0]), dt)
    for k in range(K):
        const = dt.type(-0.5 * D * LOG_2PI) - np.sum(np.log(stds[k]))
        out[k] = const - 0.5 * np.sum(np.square((1.0 / stds[k])[None, :] * (means[k][None, :] - X)), axis=1)
    return out


def component_log_densities(gmm: OracleGMM, X):
    if gmm.diagonal_covs:
        return diag_component_log_densities(X, gmm.means, gmm.chol_cov)
    return full_component_log_densities(X, gmm.means, gmm.chol_cov)


def component_marginal_log_densities(gmm: OracleGMM, X, dim):
    """models/full_cov_gmm.py:49-54."""
    covs = gmm.chol_cov @ np.transpose(gmm.chol_cov, (0, 2, 1))
    diffs = X[:, dim][None, :] - gmm.means[:, dim][:, None]
    mahalas = -0.5 * diffs * diffs / covs[:, dim, dim][:, None]
    const = -0.5 * np.log(covs[:, dim, dim]) - gmm.dt.type(0.5 * LOG_2PI)
    return mahalas + const[:, None]


def gaussian_entropy(gmm: OracleGMM, chol):
    """full_cov_gmm.py:33-34 / diagonal_gmm.py:40-41."""
    D = gmm.num_dimensions
    diag = chol if gmm.diagonal_covs else np.diag(chol)
    return gmm.dt.type(0.5 * D * (LOG_2PI + 1)) + np.sum(np.log(diag))


def component_entropies(gmm: OracleGMM):
    return np.array([gaussian_entropy(gmm, gmm.chol_cov[k]) for k in range(gmm.num_components)], gmm.dt)


def get_average_entropy(gmm: OracleGMM):
    """gmm.py:262-272."""
    return np.sum(gmm.weights * component_entropies(gmm))


# --------------------------------------------------------------------------------------
# A2 / A3: mixture density and gradient
# --------------------------------------------------------------------------------------
def log_densities_also_individual(gmm: OracleGMM, X):
    """models/gmm.py:183-201."""
    lq = component_log_densities(gmm, X)
    return logsumexp(lq + gmm.log_weights[:, None], axis=0), lq


def log_density(gmm: OracleGMM, X):
    """models/gmm.py:203-216."""
    return log_densities_also_individual(gmm, X)[0]


def log_density_and_grad(gmm: OracleGMM, X):
    """models/gmm.py:274-300.  The reference differentiates through the triangular solve with a
    GradientTape; restated analytically: grad = -sum_k r_kn Sigma_k^{-1}(x_n - mu_k), r = softmax_k."""
    lqmix, lq = log_densities_also_individual(gmm, X)
    resp = np.exp(lq + gmm.log_weights[:, None] - lqmix[None, :])
    grad = np.zeros_like(X)
    for k in range(gmm.num_components):
        diffs = X - gmm.means[k]
        if gmm.diagonal_covs:
            ptd = diffs / np.square(gmm.chol_cov[k])[None, :]
        else:
            z = solve_triangular(gmm.chol_cov[k], diffs.T, lower=True, check_finite=False)
            ptd = solve_triangular(gmm.chol_cov[k], z, lower=True, trans="T", check_finite=False).T
        grad -= resp[k][:, None] * ptd
    return lqmix, grad.astype(X.dtype), lq


# --------------------------------------------------------------------------------------
# A5: sampling with injected noise
# --------------------------------------------------------------------------------------
def sample_from_component(gmm: OracleGMM, k, eps_Dn):
    """full_cov_gmm.py:36-39 / diagonal_gmm.py:43-45.  eps_Dn is the (D, n) standard-normal draw."""
    if gmm.diagonal_covs:
        return (gmm.means[k][:, None] + gmm.chol_cov[k][:, None] * eps_Dn).T
    return (gmm.means[k][:, None] + gmm.chol_cov[k] @ eps_Dn).T


def sample_from_components_no_shuffle(gmm: OracleGMM, n_per, noise_fn):
    """gmm.py:361-386.  noise_fn(k, D, n) -> (D, n) array replaces tf.random.normal."""
    n_per = np.asarray(n_per, np.int32)
    mapping = np.repeat(np.arange(gmm.num_components, dtype=np.int32), n_per)
    parts = [sample_from_component(gmm, k, np.asarray(noise_fn(k, gmm.num_dimensions, int(n_per[k])), gmm.dt))
             for k in range(gmm.num_components)]
    return np.concatenate(parts, axis=0).astype(gmm.dt), mapping


def sample_categorical(gmm: OracleGMM, u):
    """gmm.py:124-137 with the uniform draw u[n] injected; all-False argmax -> 0 (quirk 5)."""
    thresholds = np.cumsum(gmm.weights)[None, :]
    return np.argmax(u[:, None] < thresholds, axis=-1).astype(np.int32)


# --------------------------------------------------------------------------------------
# A6 / A7: sample database
# --------------------------------------------------------------------------------------
class OracleSampleDB:
    """optimization/sample_db.py."""

    def __init__(self, dim, diagonal_covariances, keep_samples, max_samples=None, dt=np.float32):
        self._dim, self.diagonal_covariances = dim, diagonal_covariances
        self.keep_samples, self.max_samples, self.dt = keep_samples, max_samples, np.dtype(dt)
        cshape = (0, dim) if diagonal_covariances else (0, dim, dim)
        self.samples = np.zeros((0, dim), dt)
        self.means = np.zeros((0, dim), dt)
        self.chols = np.zeros(cshape, dt)
        self.inv_chols = np.zeros(cshape, dt)
        self.target_lnpdfs = np.zeros(0, dt)
        self.target_grads = np.zeros((0, dim), dt)
        self.mapping = np.zeros(0, np.int32)
        self.num_samples_written = 0

    def remove_every_nth_sample(self, N):
        """sample_db.py:64-79."""
        self.samples = self.samples[::N]
        self.target_lnpdfs = self.target_lnpdfs[::N]
        self.target_grads = self.target_grads[::N]
        self.mapping = self.mapping[::N]
        used, reduced, _ = unique_with_counts_first_occurrence(self.mapping)
        self.mapping = reduced
        self.means, self.chols, self.inv_chols = self.means[used], self.chols[used], self.inv_chols[used]

    def _inv(self, chols):
        if self.diagonal_covariances:
            return (1.0 / chols).astype(self.dt)
        return np.stack([np.linalg.inv(c) for c in chols]).astype(self.dt) if len(chols) else chols.copy()

    def add_samples(self, samples, means, chols, target_lnpdfs, target_grads, mapping):
        """sample_db.py:82-135."""
        if self.max_samples is not None and samples.shape[0] + self.samples.shape[0] > self.max_samples:
            self.remove_every_nth_sample(2)
        self.num_samples_written += samples.shape[0]
        if self.keep_samples:
            self.mapping = np.concatenate((self.mapping, mapping + self.chols.shape[0])).astype(np.int32)
            self.means = np.concatenate((self.means, means))
            self.chols = np.concatenate((self.chols, chols))
            self.inv_chols = np.concatenate((self.inv_chols, self._inv(chols)))
            self.samples = np.concatenate((self.samples, samples))
            self.target_lnpdfs = np.concatenate((self.target_lnpdfs, target_lnpdfs))
            self.target_grads = np.concatenate((self.target_grads, target_grads))
        else:
            self.mapping = np.asarray(mapping, np.int32)
            self.means, self.chols, self.inv_chols = means.copy(), chols.copy(), self._inv(chols)
            self.samples, self.target_lnpdfs, self.target_grads = samples, target_lnpdfs, target_grads

    def gaussian_log_pdf(self, mean, chol, inv_chol, x):
        """sample_db.py:154-162 (dense inv_chol matmul, not a triangular solve)."""
        dt = self.dt
        if self.diagonal_covariances:
            const = dt.type(-0.5 * self._dim * LOG_2PI) - np.sum(np.log(chol))
            return const - 0.5 * np.sum(np.square(inv_chol[:, None] * (mean[None, :] - x).T), axis=0)
        const = dt.type(-0.5 * self._dim * LOG_2PI) - np.sum(np.log(np.diag(chol)))
        return const - 0.5 * np.sum(np.square(inv_chol @ (mean - x).T), axis=0)

    def evaluate_background(self, weights, means, chols, inv_chols, samples):
        """sample_db.py:164-192: sequential pairwise logsumexp."""
        with np.errstate(divide="ignore"):
            log_weights = np.log(weights)
        log_pdfs = self.gaussian_log_pdf(means[0], chols[0], inv_chols[0], samples) + log_weights[0]
        for i in range(1, len(weights)):
            nxt = self.gaussian_log_pdf(means[i], chols[i], inv_chols[i], samples) + log_weights[i]
            log_pdfs = logsumexp(np.stack((log_pdfs, nxt), axis=0), axis=0)
        return log_pdfs.astype(self.dt)

    def get_random_sample(self, N, shuffle_fn):
        """sample_db.py:137-152: the first N entries of a random permutation of the database (shuffle_fn(n) -> the
        permutation tf.random.shuffle(tf.range(n)) produced)."""
        chosen = np.asarray(shuffle_fn(self.samples.shape[0]), np.int64)[:int(N)]
        return self.samples[chosen], self.target_lnpdfs[chosen]

    def get_newest_samples(self, N):
        """sample_db.py:195-228 -> (bg, samples, mapping, lnpdfs, grads)."""
        dt, D = self.dt, self._dim
        if self.samples.shape[0] == 0 or N == 0:
            return np.zeros(0, dt), np.zeros((0, D), dt), np.zeros(0, np.int32), np.zeros(0, dt), np.zeros((0, D), dt)
        start = max(0, self.samples.shape[0] - int(N))
        X, lnpdfs, grads, amap = self.samples[start:], self.target_lnpdfs[start:], self.target_grads[start:], self.mapping[start:]
        comps, _, count = unique_with_counts_first_occurrence(amap)
        count = count.astype(dt)
        weight = count / np.sum(count)
        bg = self.evaluate_background(weight, self.means[comps], self.chols[comps], self.inv_chols[comps], X)
        return bg, X, amap, lnpdfs, grads


# --------------------------------------------------------------------------------------
# A8: VIPS sample selection
# --------------------------------------------------------------------------------------
def get_effective_samples(model_densities, oldsamples_pdf):
    """gmmvi_modules/sample_selector.py:140-158."""
    lw = model_densities - oldsamples_pdf[None, :]
    lw = lw - logsumexp(lw, axis=1, keepdims=True)
    w = np.exp(lw)
    return 1.0 / np.sum(w * w, axis=1)


def vips_num_additional_samples(gmm: OracleGMM, old_samples, old_bg, desired):
    """sample_selector.py:186-199 -> int32 [K]."""
    if old_samples.shape[0] == 0:
        n_eff = np.zeros(gmm.num_components, np.int32)
    else:
        n_eff = np.floor(get_effective_samples(component_log_densities(gmm, old_samples), old_bg)).astype(np.int32)
    return np.maximum(1, desired - n_eff).astype(np.int32)


def vips_select_samples(gmm: OracleGMM, db: OracleSampleDB, target, desired, ratio_reused, noise_fn):
    """sample_selector.py:204-219 -> (samples, mapping, bg, lnpdfs, grads)  (quirk 2)."""
    reused = int(math.floor(ratio_reused * desired))
    old_bg, old_X, _, _, _ = db.get_newest_samples(reused * gmm.num_components)
    n_reused = old_X.shape[0]
    n_add = vips_num_additional_samples(gmm, old_X, old_bg, desired)
    new_X, mapping = sample_from_components_no_shuffle(gmm, n_add, noise_fn)
    new_lnpdf, new_grad = target(new_X)
    db.add_samples(new_X, gmm.means, gmm.chol_cov, new_lnpdf, new_grad, mapping)
    bg, X, mapping, lnpdfs, grads = db.get_newest_samples(n_reused + new_X.shape[0])
    return X, mapping, bg, lnpdfs, grads


def gmm_sample(gmm: OracleGMM, u, noise_fn):
    """GMM.sample (gmm.py:139-163) with the uniform draws u[n] and the normal draws injected: the samples come back
    GROUPED BY COMPONENT, the component indices in DRAW ORDER (quirk 4)."""
    comps = sample_categorical(gmm, np.asarray(u, gmm.dt))
    counts = np.bincount(comps, minlength=gmm.num_components)
    parts = [sample_from_component(gmm, k, np.asarray(noise_fn(k, gmm.num_dimensions, int(counts[k])), gmm.dt))
             for k in range(gmm.num_components)]
    return np.concatenate(parts, axis=0).astype(gmm.dt), comps


def lin_select_samples(gmm: OracleGMM, db: OracleSampleDB, target, desired, ratio_reused, uniform_fn, noise_fn):
    """LinSampleSelector.select_samples (sample_selector.py:258-339): `desired` is the TOTAL number of samples, the
    effective sample size is the mixture's (:273-276), new samples are drawn from the mixture with GMM.sample, whose
    mapping is not aligned with its samples (quirk 4) - stored in the database as is.  uniform_fn(n) -> u[n]."""
    reused = int(math.floor(ratio_reused * desired))
    old_bg, old_X, _, _, _ = db.get_newest_samples(reused * gmm.num_components)
    n_reused = old_X.shape[0]
    if n_reused == 0:
        n_eff = 0
    else:
        n_eff = int(np.floor(get_effective_samples(log_density(gmm, old_X)[None, :], old_bg))[0])
    n_add = max(1, desired - n_eff)
    new_X, mapping = gmm_sample(gmm, uniform_fn(n_add), noise_fn)
    new_lnpdf, new_grad = target(new_X)
    db.add_samples(new_X, gmm.means, gmm.chol_cov, new_lnpdf, new_grad, mapping)
    bg, X, mapping, lnpdfs, grads = db.get_newest_samples(n_reused + new_X.shape[0])
    return X, mapping, bg, lnpdfs, grads


# --------------------------------------------------------------------------------------
# A9: Stein natural-gradient estimator
# --------------------------------------------------------------------------------------
def _stable_expectation(log_weights, values):
    """ng_estimator.py:147-152."""
    n = log_weights.dtype.type(log_weights.shape[0])
    lw = log_weights.reshape(log_weights.shape + (1,) * (values.ndim - log_weights.ndim))
    with np.errstate(divide="ignore"):
        logx = lw + np.log(np.abs(values))
    lswe, signs = reduce_weighted_logsumexp_sign(logx, np.sign(values), axis=0)
    return 1 / n * signs * np.exp(lswe)


def stein_for_comp(chol, mean, comp_lq, X, bg, G, diagonal, self_normalized):
    """ng_estimator.py:154-188 -> (expected_gradient, expected_hessian)."""
    if self_normalized:                                                   # :171-188
        lw = comp_lq - bg
        lw = lw - logsumexp(lw, axis=0, keepdims=True)
        w = np.exp(lw)
        iw = w / np.sum(w, axis=0, keepdims=True)
        WG = iw[:, None] * G
        if diagonal:
            ptd = (1 / (chol ** 2))[:, None] * (X - mean).T
            H = np.sum(ptd.T * WG, axis=0)
        else:
            z = solve_triangular(chol, (X - mean).T, lower=True, check_finite=False)
            ptd = solve_triangular(chol, z, lower=True, trans="T", check_finite=False)     # cholesky_solve
            H = WG.T @ ptd.T                                              # H[i,j] = sum_n WG[n,i] ptd[j,n]
            H = 0.5 * (H + H.T)
        return np.sum(WG, axis=0), H
    lw = comp_lq - bg                                                     # :154-169
    eg = _stable_expectation(lw, G)
    if diagonal:
        ptd = (1 / (chol ** 2))[:, None] * (X - mean).T
        H = _stable_expectation(lw, ptd.T * G)
    else:
        z = solve_triangular(chol, (X - mean).T, lower=True, check_finite=False)
        ptd = solve_triangular(chol, z, lower=True, trans="T", check_finite=False)
        H = _stable_expectation(lw, ptd.T[:, None, :] * G[:, :, None])    # not symmetrised (quirk 7)
    return eg, H


def _rewards_for_comp(i, only_own, X, rel_map, lq, log_ratios, G, bg):
    """ng_estimator.py:107-120."""
    if only_own:
        idx = np.where(rel_map == i)[0]
        return X[idx], log_ratios[idx], (G[idx] if G is not None else None), lq[i][idx], lq[i][idx]
    return X, log_ratios, G, bg, lq[i]


def stein_ng(gmm: OracleGMM, X, mapping, bg, target_lnpdfs, target_grads,
             only_use_own_samples=False, use_self_normalized_importance_weights=True):
    """ng_estimator.py:204-263 -> (expected_hessian_neg, expected_gradient_neg)."""
    K = gmm.num_components
    rel_map = mapping - np.max(mapping) + K - 1                              # :244
    lqmix, grad_q, lq = log_density_and_grad(gmm, X)
    log_ratios = target_lnpdfs - lqmix
    G = target_grads - grad_q
    Hs, gs = [], []
    for i in range(K):
        mX, _, mG, mbg, mlq = _rewards_for_comp(i, only_use_own_samples, X, rel_map, lq, log_ratios, G, bg)
        eg, H = stein_for_comp(gmm.chol_cov[i], gmm.means[i], mlq, mX, mbg, mG, gmm.diagonal_covs,
                               use_self_normalized_importance_weights)
        Hs.append(-H)
        gs.append(-eg)
    return np.stack(Hs).astype(gmm.dt), np.stack(gs).astype(gmm.dt)


# --------------------------------------------------------------------------------------
# A10: MORE natural-gradient estimator
# --------------------------------------------------------------------------------------
def quad_features(x):
    """optimization/least_squares.py:113-124."""
    n, D = x.shape
    quad = [x[:, i:i + 1] * x[:, i:] for i in range(D)]
    return np.concatenate(quad + [x, np.ones((n, 1), x.dtype)], axis=1)


def fit_quadratic(regularizer, X, y, weights, mean, chol):
    """least_squares.py:126-191 (+ RegressionFunc.fit :34-76) -> (quad_term, lin_term, const_term)."""
    dt = X.dtype
    D = X.shape[1]
    inv_chol = np.linalg.inv(chol).astype(dt)
    Z = (X - mean) @ inv_chol.T
    Phi = quad_features(Z)
    F = Phi.shape[1]
    WPhiT = (weights[:, None] * Phi).T
    reg = np.eye(F, dtype=dt) * dt.type(regularizer)
    reg[F - 1, F - 1] = 0                                                    # bias entry, :69-73
    params = np.linalg.solve(WPhiT @ Phi + reg, WPhiT @ y[:, None])[:, 0].astype(dt)
    qt = np.zeros((D, D), dt)
    qt[np.triu_indices(D)] = params[:-(D + 1)]
    quad = -qt - qt.T
    lin = params[-(D + 1):-1]
    const = params[-1]
    quad = inv_chol.T @ quad @ inv_chol
    t1 = inv_chol.T @ lin
    t2 = quad @ mean
    lin = t1 + t2
    const = const + np.sum(mean * (-0.5 * t2 - t1))
    return quad.astype(dt), lin.astype(dt), dt.type(const)


def more_ng(gmm: OracleGMM, X, mapping, bg, target_lnpdfs, target_grads=None,
            only_use_own_samples=False, use_self_normalized_importance_weights=True, **_):
    """ng_estimator.py:296-376."""
    K = gmm.num_components
    rel_map = mapping - np.max(mapping) + K - 1
    lqmix, lq = log_densities_also_individual(gmm, X)
    log_ratios = target_lnpdfs - lqmix
    Hs, gs = [], []
    for i in range(K):
        mX, my, _, mbg, mlq = _rewards_for_comp(i, only_use_own_samples, X, rel_map, lq, log_ratios, None, bg)
        lw = mlq - mbg
        if use_self_normalized_importance_weights:
            lw = lw - logsumexp(lw, axis=0, keepdims=True)
            w = np.exp(lw)
            iw = w / np.sum(w, axis=0, keepdims=True)
        else:
            iw = np.exp(lw)
        quad, lin, _ = fit_quadratic(gmm.l2_regularizers[i], mX, my, iw.astype(gmm.dt), gmm.means[i], gmm.chol_cov[i])
        Hs.append(quad)
        gs.append(quad @ gmm.means[i] - lin)
    return np.stack(Hs).astype(gmm.dt), np.stack(gs).astype(gmm.dt)


# --------------------------------------------------------------------------------------
# A11 / A12: component updaters
# --------------------------------------------------------------------------------------
def _l2_rule(gmm: OracleGMM, successes):
    """ng_based_component_updater.py:135-138 (quirk 11)."""
    dt = gmm.dt
    gmm.l2_regularizers = np.where(successes,
                                   np.maximum(dt.type(0.5) * gmm.l2_regularizers, dt.type(gmm.initial_regularizer)),
                                   np.minimum(dt.type(1e-6), dt.type(10) * gmm.l2_regularizers)).astype(dt)


def kl_eval(eta, old_lin, old_prec, old_inv_chol, reward_lin, reward_quad, kl_const, old_mean, diagonal,
            eta_in_logspace):
    """ng_based_component_updater.py:244-333 -> (kl, new_mean, new_precision, inv_chol_inv)."""
    dt = old_mean.dtype
    D = old_mean.shape[0]
    if eta_in_logspace:
        eta = np.exp(eta)
    eta = dt.type(eta)
    new_lin = (eta * old_lin + reward_lin) / eta
    new_prec = (eta * old_prec + reward_quad) / eta
    if diagonal:
        with np.errstate(invalid="ignore", divide="ignore"):
            chol_prec = np.sqrt(new_prec)
            new_mean = 1.0 / new_prec * new_lin
            ici = 1.0 / chol_prec
            diff = old_mean - new_mean
            kl = 0.5 * (np.maximum(dt.type(0), np.sum(np.log(new_prec / old_prec) + old_prec / new_prec) - D)
                        + np.sum(np.square(old_inv_chol * diff)))
        return dt.type(kl), new_mean, new_prec, ici
    chol_prec = cholesky_or_nan(new_prec)
    if np.any(np.isnan(chol_prec)):
        return dt.type(FLT_MAX), old_mean, old_prec, old_inv_chol
    y = solve_triangular(chol_prec, new_lin, lower=True, check_finite=False)
    new_mean = solve_triangular(chol_prec, y, lower=True, trans="T", check_finite=False).astype(dt)
    new_logdet = -2 * np.sum(np.log(np.diag(chol_prec)))
    ici = np.linalg.inv(chol_prec).astype(dt)
    trace_term = np.square(np.linalg.norm(ici @ old_inv_chol.T))
    diff = old_mean - new_mean
    kl = 0.5 * (kl_const - new_logdet + trace_term + np.sum(np.square(old_inv_chol @ diff)))
    return dt.type(kl), new_mean, new_prec.astype(dt), ici


def kl_bracketing_search(kl_bound, lower, upper, args, diagonal, trace: Optional[list] = None):
    """ng_based_component_updater.py:335-429 (always called with eta_in_logspace=True, :473-475)."""
    dt = args[-1].dtype
    lower, upper, kl_bound = dt.type(lower), dt.type(upper), dt.type(kl_bound)
    eta = dt.type(0.5) * (upper + lower)
    feasible = False
    for _ in range(1000):
        diff = min(np.exp(upper) - np.exp(eta), np.exp(eta) - np.exp(lower))
        if diff < 1e-1:
            break
        kl = kl_eval(eta, *args, diagonal, True)[0]
        if trace is not None:
            trace.append((float(eta), float(kl)))
        if abs(kl_bound - kl) < dt.type(1e-1) * kl_bound:
            lower = upper = eta
            break
        if kl_bound > kl:
            upper = eta
            feasible = True
        else:
            lower = eta
        eta = dt.type(0.5) * (upper + lower)
    if feasible:
        lower = upper
    return np.exp(lower), np.exp(upper)


def kl_constrained_update(gmm: OracleGMM, H_neg, g_neg, stepsizes, temperature=1.0, traces: Optional[list] = None):
    """KLConstrainedNgBasedComponentUpdater.apply_NG_update, ng_based_component_updater.py:431-524.
    Returns dict(success, etas, kls)."""
    dt, D, K = gmm.dt, gmm.num_dimensions, gmm.num_components
    means, chols, succ, kls, etas = [], [], [], [], []
    for i in range(K):
        old_chol, old_mean, last_eta, eps = gmm.chol_cov[i], gmm.means[i], gmm.last_log_etas[i], stepsizes[i]
        R = H_neg[i]
        if gmm.diagonal_covs:                                              # :447-453
            r_lin = R * old_mean - g_neg[i]
            old_logdet = 2 * np.sum(np.log(old_chol))
            old_inv_chol = 1.0 / old_chol
            old_prec = old_inv_chol ** 2
            old_lin = old_prec * old_mean
        else:                                                              # :454-460
            r_lin = R @ old_mean - g_neg[i]
            old_logdet = 2 * np.sum(np.log(np.diag(old_chol)))
            old_inv_chol = np.linalg.inv(old_chol).astype(dt)
            old_prec = old_inv_chol.T @ old_inv_chol
            old_lin = old_prec @ old_mean
        kl_const = dt.type(old_logdet - D)
        if last_eta < 0:                                                   # :462-471
            lb, ub = dt.type(-20.0), dt.type(80.0)
        else:
            lb = max(dt.type(0.0), np.log(last_eta) - dt.type(3))
            ub = np.log(last_eta) + dt.type(3)
        args = (old_lin.astype(dt), old_prec.astype(dt), old_inv_chol, r_lin.astype(dt), R, kl_const, old_mean)
        tr = [] if traces is not None else None
        new_lower, new_upper = kl_bracketing_search(eps, lb, ub, args, gmm.diagonal_covs, tr)
        if traces is not None:
            traces.append(tr)
        eta = max(dt.type(new_lower), dt.type(temperature))               # :476
        success = False
        if new_lower == new_upper:                                         # :478
            success = True
            kl, new_mean, _, ici = kl_eval(eta, *args, gmm.diagonal_covs, False)
            new_cov = np.square(ici) if gmm.diagonal_covs else ici.T @ ici
            if kl < FLT_MAX:
                if gmm.diagonal_covs:
                    new_chol = np.sqrt(new_cov)
                else:
                    new_chol = cholesky_or_nan(new_cov)
                    if np.any(np.isnan(new_chol)):
                        success = False
            else:
                success = False
            if gmm.diagonal_covs and success and not np.all(np.isfinite(new_chol)):
                pass  # the reference does not test diagonal NaNs (:488-489); keep its behaviour
        if success:
            chols.append(new_chol.astype(dt)); means.append(new_mean.astype(dt))
            kls.append(dt.type(kl)); etas.append(dt.type(eta))
        else:
            chols.append(old_chol); means.append(old_mean); kls.append(dt.type(-1)); etas.append(dt.type(-1))
        succ.append(success)
    succ = np.array(succ)
    gmm.means, gmm.chol_cov = np.stack(means), np.stack(chols)          # replace_components, gmm.py:401-418
    gmm.num_received_updates = gmm.num_received_updates + 1
    _l2_rule(gmm, succ)
    gmm.last_log_etas = np.array(etas, dt)                                # stores eta, not log eta (quirk 10)
    return dict(success=succ, etas=np.array(etas, dt), kls=np.array(kls, dt))


def direct_update(gmm: OracleGMM, H_neg, g_neg, stepsizes):
    """DirectNgBasedComponentUpdater.apply_NG_update, ng_based_component_updater.py:97-141 (full cov only)."""
    dt, K = gmm.dt, gmm.num_components
    means, chols, succ = [], [], []
    for i in range(K):
        old_chol, old_mean, s = gmm.chol_cov[i], gmm.means[i], stepsizes[i]
        iL = np.linalg.inv(old_chol).astype(dt)
        P = iL.T @ iL
        old_lin = P @ old_mean
        d_lin = H_neg[i] @ old_mean - g_neg[i]
        new_lin = old_lin + s * d_lin
        new_P = P + s * H_neg[i]
        try:
            new_mean = np.linalg.solve(new_P, new_lin).astype(dt)
            new_cov = np.linalg.inv(new_P).astype(dt)
            new_chol = cholesky_or_nan(new_cov)
        except np.linalg.LinAlgError:
            new_chol = np.full_like(old_chol, np.nan)
        if np.any(np.isnan(new_chol)):
            succ.append(False); means.append(old_mean); chols.append(old_chol)
        else:
            succ.append(True); means.append(new_mean); chols.append(new_chol.astype(dt))
    succ = np.array(succ)
    _l2_rule(gmm, succ)
    gmm.means, gmm.chol_cov = np.stack(means), np.stack(chols)
    gmm.num_received_updates = gmm.num_received_updates + 1
    return dict(success=succ)


def iblr_update(gmm: OracleGMM, H_neg, g_neg, stepsizes):
    """NgBasedComponentUpdaterIblr.apply_NG_update, ng_based_component_updater.py:160-223."""
    dt, K = gmm.dt, gmm.num_components
    means, chols, succ = [], [], []
    for i in range(K):
        old_chol, old_mean, s, R = gmm.chol_cov[i], gmm.means[i], stepsizes[i], H_neg[i]
        if gmm.diagonal_covs:
            corr = s / 2 * R * old_chol * old_chol * R
            iL = 1.0 / old_chol
            P = iL * iL
        else:
            corr = s / 2 * R @ old_chol @ old_chol.T @ R
            iL = np.linalg.inv(old_chol).astype(dt)
            P = iL.T @ iL
        d_prec = R + corr
        d_mean = -g_neg[i]
        if gmm.num_received_updates[i] == 0:                                # quirk 12
            new_mean = old_mean
        elif gmm.diagonal_covs:
            new_mean = old_mean + s * old_chol * old_chol * d_mean
        else:
            new_mean = old_mean + s * old_chol @ old_chol.T @ d_mean
        new_P = P + s * d_prec
        with np.errstate(invalid="ignore", divide="ignore"):
            if gmm.diagonal_covs:
                new_chol = np.sqrt(1.0 / new_P)
            else:
                try:
                    new_chol = cholesky_or_nan(np.linalg.inv(new_P).astype(dt))
                except np.linalg.LinAlgError:
                    new_chol = np.full_like(old_chol, np.nan)
        if np.any(np.isnan(new_chol)):
            succ.append(False); means.append(old_mean); chols.append(old_chol)
        else:
            succ.append(True); means.append(new_mean.astype(dt)); chols.append(new_chol.astype(dt))
    succ = np.array(succ)
    _l2_rule(gmm, succ)
    gmm.means, gmm.chol_cov = np.stack(means), np.stack(chols)
    gmm.num_received_updates = gmm.num_received_updates + 1
    return dict(success=succ)


# --------------------------------------------------------------------------------------
# A13 / A14: weight update
# --------------------------------------------------------------------------------------
def expected_log_ratios(gmm: OracleGMM, X, bg, target_lnpdfs, temperature=1.0, self_normalized=True):
    """gmmvi_modules/weight_updater.py:56-75 (also stores the rewards)."""
    dt = gmm.dt
    lqmix, lq = log_densities_also_individual(gmm, X)
    log_ratios = target_lnpdfs - dt.type(temperature) * lqmix
    if self_normalized:
        lw = lq - bg
        lw = lw - logsumexp(lw, axis=1, keepdims=True)
        w = np.exp(lw)
        iw = w / np.sum(w, axis=1, keepdims=True)
        elr = iw @ log_ratios
    else:
        n = dt.type(X.shape[0])
        lw = lq - bg
        with np.errstate(divide="ignore"):
            lswe, signs = reduce_weighted_logsumexp_sign(lw + np.log(np.abs(log_ratios)), np.sign(log_ratios), axis=1)
        elr = 1 / n * signs * np.exp(lswe)
    elr = elr.astype(dt)
    gmm.store_rewards(dt.type(temperature) * gmm.log_weights + elr)
    return elr


def direct_weight_update(gmm: OracleGMM, elr, stepsize, temperature=1.0):
    """weight_updater.py:123-141."""
    dt = gmm.dt
    if gmm.num_components > 1:
        v = gmm.log_weights + dt.type(stepsize) / dt.type(temperature) * elr
        v = v - logsumexp(v)
        v = np.maximum(v, dt.type(-69.07))
        v = v - logsumexp(v)
        gmm.wrapper_replace_weights(v)


def weight_kl(gmm: OracleGMM, eta, rewards, temperature):
    """weight_updater.py:164-191."""
    dt = gmm.dt
    eta, T = dt.type(eta), dt.type(temperature)
    v = (eta + 1) / (T + eta) * gmm.log_weights + dt.type(1.0) / (T + eta) * rewards
    nl = v - logsumexp(v)
    nl = np.maximum(nl, dt.type(-69.07))
    nl = nl - logsumexp(nl)
    kl = np.sum(np.exp(nl) * (nl - gmm.log_weights))
    return dt.type(kl), nl.astype(dt)


def trust_region_weight_update(gmm: OracleGMM, elr, kl_bound, temperature=1.0, trace: Optional[list] = None):
    """weight_updater.py:193-279 -> (kl, eta)."""
    dt = gmm.dt
    if gmm.num_components <= 1:
        return dt.type(-1), dt.type(-1)
    lower, upper, kl_bound = dt.type(-45.0), dt.type(45.0), dt.type(kl_bound)
    log_eta = dt.type(0.5) * (upper + lower)
    feasible = False
    kl, eta, new_lw = dt.type(-1), dt.type(-1), gmm.log_weights
    for _ in range(50):
        eta = np.exp(log_eta)
        diff = abs(np.exp(upper) - np.exp(lower))
        if diff < 1e-1:
            break
        kl, new_lw = weight_kl(gmm, eta, elr, temperature)
        if trace is not None:
            trace.append((float(log_eta), float(kl)))
        if abs(kl_bound - kl) < dt.type(1e-1) * kl_bound:
            lower = upper
            break
        if kl_bound > kl:
            upper = log_eta
            feasible = True
        else:
            lower = log_eta
        log_eta = dt.type(0.5) * (upper + lower)
    if lower == upper:
        pass
    elif feasible:
        kl, new_lw = weight_kl(gmm, np.exp(upper), elr, temperature)
        eta = np.exp(upper)
    else:
        kl, eta, new_lw = dt.type(-1), dt.type(-1), gmm.log_weights
    gmm.wrapper_replace_weights(new_lw)
    return kl, eta


# --------------------------------------------------------------------------------------
# stepsize glue (A.12)
# --------------------------------------------------------------------------------------
def improvement_based_component_stepsize(gmm: OracleGMM, min_stepsize, max_stepsize, inc, dec):
    """gmmvi_modules/component_stepsize_adaptation.py:165-188 (quirk 14)."""
    dt = gmm.dt
    h = gmm.reward_history
    worse = h[:, -2] >= h[:, -1]
    return np.where(worse, np.maximum(dt.type(dec) * gmm.stepsizes, dt.type(min_stepsize)),
                    np.minimum(dt.type(inc) * gmm.stepsizes, dt.type(max_stepsize))).astype(dt)


def decaying_component_stepsize(gmm: OracleGMM, initial_stepsize, annealing_exponent):
    """component_stepsize_adaptation.py:116-130."""
    dt = gmm.dt
    return (dt.type(initial_stepsize) / (1 + np.power(gmm.num_received_updates, dt.type(annealing_exponent)))).astype(dt)


class DecayingWeightStepsize:
    """weight_stepsize_adaptation.py:70-106: initial / (1 + n^exponent), n = number of previous weight updates."""

    def __init__(self, initial, annealing_exponent, dt=np.float32):
        self.dt = np.dtype(dt)
        self.initial, self.exponent, self.n = self.dt.type(initial), annealing_exponent, 0.0

    def update(self, gmm: "OracleGMM"):
        stepsize = self.initial / (1.0 + self.dt.type(self.n) ** self.exponent)
        self.n += 1.0
        return self.dt.type(stepsize)


class ImprovementBasedWeightStepsize:
    """gmmvi_modules/weight_stepsize_adaptation.py:108-156."""

    def __init__(self, initial, min_stepsize, max_stepsize, inc, dec, dt=np.float32):
        dt = np.dtype(dt)
        self.dt, self.stepsize = dt, dt.type(initial)
        self.min, self.max, self.inc, self.dec = (dt.type(v) for v in (min_stepsize, max_stepsize, inc, dec))
        self.elbo_history = [dt.type(FLT_MIN)]

    def update(self, gmm: OracleGMM):
        elbo = np.sum(gmm.weights * gmm.reward_history[:, -1]) - np.sum(gmm.weights * gmm.log_weights)
        self.elbo_history.append(self.dt.type(elbo))
        if self.elbo_history[-1] > self.elbo_history[-2]:
            self.stepsize = min(self.inc * self.stepsize, self.max)
        else:
            self.stepsize = max(self.dec * self.stepsize, self.min)
        return self.stepsize


# --------------------------------------------------------------------------------------
# Targets used as fixtures (experiments/target_distributions/gmm.py:28-40)
# --------------------------------------------------------------------------------------
def gmm_target(weights, means, covs, dt=np.float32):
    """GMM_LNPDF.log_density + analytic gradient (the reference uses a GradientTape,
    gmmvi_modules/sample_selector.py:69-78).  Returns f(X) -> (lnpdf[N], grad[N, D])."""
    tgt = make_full_gmm(weights, means, covs, dt)

    def f(X):
        lq, g, _ = log_density_and_grad(tgt, X)
        return lq, g
    return f


def student_t_mixture_target(weights, means, covs, alpha=2, dt=np.float64):
    """StudentTMixture_LNPDF (experiments/target_distributions/student_t_mixture.py:34-68): MixtureSameFamily of
    MultivariateStudentTLinearOperator(df=alpha, loc=means, scale=chol(covs)),
        log t_j(x) = lgamma((nu+D)/2) - lgamma(nu/2) - D/2 log(nu pi) - sum_i log L_j[i,i] - (nu+D)/2 log1p(m_j(x)/nu),
        m_j(x) = |L_j^-1 (x - mu_j)|^2,   log p(x) = LSE_j(log w_j + log t_j(x)).
    The reference differentiates with a GradientTape (use_log_density_and_grad=False, sample_selector.py:73-77); the
    analytic gradient is  -sum_j r_j(x) (nu+D)/(nu+m_j) Sigma_j^-1 (x - mu_j).  Returns f(X) -> (lnpdf[N], grad[N, D])."""
    from math import lgamma
    w = np.asarray(weights, dt)
    means = np.asarray(means, dt)
    chols = np.stack([cholesky_or_nan(np.asarray(c, dt)) for c in np.asarray(covs, dt)])
    D = means.shape[1]
    nu = float(alpha)
    norm = lgamma(0.5 * (nu + D)) - lgamma(0.5 * nu) - 0.5 * D * np.log(nu * np.pi)
    logdet = np.sum(np.log(np.diagonal(chols, axis1=1, axis2=2)), axis=1)
    import scipy.linalg as sla

    def f(X):
        X = np.asarray(X, dt)
        lt = np.empty((means.shape[0], X.shape[0]), dt)
        pz = np.empty((means.shape[0],) + X.shape, dt)                  # Sigma_j^-1 (x - mu_j)
        maha = np.empty_like(lt)
        for j in range(means.shape[0]):
            diff = (X - means[j]).T
            z = sla.solve_triangular(chols[j], diff, lower=True)
            maha[j] = np.sum(z * z, axis=0)
            pz[j] = sla.solve_triangular(chols[j], z, lower=True, trans="T").T
            lt[j] = norm - logdet[j] - 0.5 * (nu + D) * np.log1p(maha[j] / nu)
        lw = lt + np.log(w)[:, None]
        lp = logsumexp(lw, axis=0)
        r = np.exp(lw - lp[None, :]) * (nu + D) / (nu + maha)
        grad = -np.einsum("jn,jnd->nd", r, pz)
        return lp.astype(dt), grad.astype(dt)
    return f


def student_t_mixture_marginal_log_density(weights, means, covs, X, dim, alpha=2, dt=np.float64):
    """StudentTMixture_LNPDF.marginal_log_density (student_t_mixture.py:46-64): mixture of scalar Student-t with
    loc = means[:, dim], scale = sqrt(covs[:, dim, dim])."""
    from math import lgamma
    nu = float(alpha)
    loc = np.asarray(means, dt)[:, dim]
    scale = np.sqrt(np.asarray(covs, dt)[:, dim, dim])
    y = (np.asarray(X, dt)[:, dim][None, :] - loc[:, None]) / scale[:, None]
    lt = (lgamma(0.5 * (nu + 1)) - lgamma(0.5 * nu) - 0.5 * np.log(nu * np.pi) - np.log(scale)[:, None]
          - 0.5 * (nu + 1) * np.log1p(y * y / nu))
    return logsumexp(lt + np.log(np.asarray(weights, dt))[:, None], axis=0)


def planar_robot_forward_kinematics(theta, link_lengths=None):
    """PlanarRobot.forward_kinematics (planar_robot.py:57-63): end effector (x, y) of the chain of joint angles."""
    theta = np.asarray(theta)
    ll = np.ones(theta.shape[1], theta.dtype) if link_lengths is None else np.asarray(link_lengths, theta.dtype)
    cs = np.cumsum(theta, axis=1)
    return np.stack((np.sum(ll * np.cos(cs), axis=1), np.sum(ll * np.sin(cs), axis=1)), axis=1)


def planar_robot_target(num_links, num_goals, prior_std=2e-1, likelihood_std=1e-2, dt=np.float64):
    """PlanarRobot (planar_robot.py:29-66): log p(theta) = N(theta; 0, diag(prior_stds^2)) + max_g N(fk(theta); goal_g,
    likelihood_std^2 I), prior_stds = [1, prior_std, ...] (:32-33), goals (7,0) | (+-7,0),(0,+-7) (:37-42), `likelihood`
    takes the MAX over the goals (:49-53).  The gradient (a GradientTape in the reference) follows the arg-max goal:
    d/dtheta_i = -theta_i / s_i^2 - [(x-gx) dx/dtheta_i + (y-gy) dy/dtheta_i] / likelihood_std^2 with
    dx/dtheta_i = -sum_{m>=i} l_m sin(c_m), dy/dtheta_i = sum_{m>=i} l_m cos(c_m).  Returns f(X) -> (lnpdf[N], grad[N, D])."""
    if num_goals == 1:
        goals = np.array([[7.0, 0.0]], dt)
    elif num_goals == 4:
        goals = np.array([[7.0, 0.0], [-7.0, 0.0], [0.0, 7.0], [0.0, -7.0]], dt)
    else:
        raise ValueError
    stds = (prior_std * np.ones(num_links)).astype(np.float32).astype(dt)      # :34 casts the scales to float32
    stds[0] = 1.0
    ls = np.asarray(likelihood_std, np.float32).astype(dt)                     # tfd.MultivariateNormalDiag(scale_diag=[..]) is fp32

    def f(X):
        X = np.asarray(X, dt)
        D = X.shape[1]
        prior = -0.5 * np.sum((X / stds) ** 2, axis=1) - np.sum(np.log(stds)) - 0.5 * D * np.log(2 * np.pi)
        cs = np.cumsum(X, axis=1)
        pos = np.stack((np.sum(np.cos(cs), axis=1), np.sum(np.sin(cs), axis=1)), axis=1)
        d = pos[None, :, :] - goals[:, None, :]                                # [G, N, 2]
        lik = -0.5 * np.sum((d / ls) ** 2, axis=2) - 2 * np.log(ls) - np.log(2 * np.pi)
        best = np.argmax(lik, axis=0)
        dbest = d[best, np.arange(X.shape[0])]                                 # [N, 2]
        dx = -np.cumsum(np.sin(cs)[:, ::-1], axis=1)[:, ::-1]
        dy = np.cumsum(np.cos(cs)[:, ::-1], axis=1)[:, ::-1]
        grad = -X / stds ** 2 - (dbest[:, :1] * dx + dbest[:, 1:] * dy) / ls ** 2
        return (prior + lik[best, np.arange(X.shape[0])]).astype(dt), grad.astype(dt)
    return f


# --------------------------------------------------------------------------------------
# A16: one full iteration with a fixed number of components (gmmvi.py:146-174)
# --------------------------------------------------------------------------------------
@dataclass
class IterationConfig:
    sample_selector: str = "component-based"        # or "mixture-based" (LinSampleSelector)
    desired_samples_per_component: int = 100
    ratio_reused_samples_to_desired: float = 0.0
    ng_estimator: str = "Stein"                     # or "MORE"
    only_use_own_samples: bool = False
    ng_self_normalized: bool = True
    updater: str = "trust-region"                   # "direct" | "iBLR"
    component_stepsize: str = "fixed"               # "improvement-based" | "decaying"
    component_stepsize_cfg: dict = field(default_factory=dict)
    weight_updater: str = "trust-region"            # "direct"
    weight_self_normalized: bool = True
    weight_stepsize: float = 1.0
    temperature: float = 1.0


def train_iter(gmm: OracleGMM, db: OracleSampleDB, target: Callable, cfg: IterationConfig, noise_fn,
               weight_stepsize_adapter: Optional[ImprovementBasedWeightStepsize] = None, uniform_fn=None):
    """GMMVI.train_iter with FixedComponentAdaptation: select -> _run_updates."""
    if cfg.sample_selector == "mixture-based":
        X, mapping, bg, lnpdfs, grads = lin_select_samples(
            gmm, db, target, cfg.desired_samples_per_component, cfg.ratio_reused_samples_to_desired, uniform_fn, noise_fn)
    else:
        X, mapping, bg, lnpdfs, grads = vips_select_samples(
            gmm, db, target, cfg.desired_samples_per_component, cfg.ratio_reused_samples_to_desired, noise_fn)
    if cfg.component_stepsize == "improvement-based":
        gmm.stepsizes = improvement_based_component_stepsize(gmm, **cfg.component_stepsize_cfg)
    elif cfg.component_stepsize == "decaying":
        gmm.stepsizes = decaying_component_stepsize(gmm, **cfg.component_stepsize_cfg)
    if cfg.ng_estimator == "Stein":
        H, g = stein_ng(gmm, X, mapping, bg, lnpdfs, grads, cfg.only_use_own_samples, cfg.ng_self_normalized)
    else:
        H, g = more_ng(gmm, X, mapping, bg, lnpdfs, grads, cfg.only_use_own_samples, cfg.ng_self_normalized)
    if cfg.updater == "trust-region":
        info = kl_constrained_update(gmm, H, g, gmm.stepsizes, cfg.temperature)
    elif cfg.updater == "direct":
        info = direct_update(gmm, H, g, gmm.stepsizes)
    else:
        info = iblr_update(gmm, H, g, gmm.stepsizes)
    ws = weight_stepsize_adapter.update(gmm) if weight_stepsize_adapter is not None else cfg.weight_stepsize
    elr = expected_log_ratios(gmm, X, bg, lnpdfs, cfg.temperature, cfg.weight_self_normalized)
    if cfg.weight_updater == "trust-region":
        trust_region_weight_update(gmm, elr, ws, cfg.temperature)
    else:
        direct_weight_update(gmm, elr, ws, cfg.temperature)
    return dict(samples=X, mapping=mapping, bg=bg, lnpdfs=lnpdfs, grads=grads, H_neg=H, g_neg=g,
                elr=elr, update=info)


# --------------------------------------------------------------------------------------
# N2: adding / deleting components (gmm_wrapper.py:90-148, component_adaptation.py:145-300)
# --------------------------------------------------------------------------------------
def add_component(gmm: OracleGMM, initial_weight, initial_mean, initial_cov):
    """GmmWrapper.add_component (gmm_wrapper.py:90-127) on top of FullCovGMM / DiagonalGMM.add_component
    (full_cov_gmm.py:64-68, diagonal_gmm.py:55-59).  The model's weights are renormalised by GMM.replace_weights, i.e.
    WITHOUT a new column in weight_history; the new component's history rows are -FLT_MAX rewards and its initial weight."""
    dt = gmm.dt
    gmm.means = np.concatenate((gmm.means, np.asarray(initial_mean, dt)[None]), axis=0)
    new_chol = np.sqrt(np.asarray(initial_cov, dt)) if gmm.diagonal_covs else cholesky_or_nan(np.asarray(initial_cov, dt))
    gmm.chol_cov = np.concatenate((gmm.chol_cov, new_chol[None]), axis=0)
    gmm.replace_weights(np.concatenate((gmm.log_weights, np.log(np.asarray([initial_weight], dt)))))
    H = gmm.reward_history.shape[1]
    gmm.l2_regularizers = np.concatenate((gmm.l2_regularizers, np.asarray([gmm.initial_regularizer], dt)))
    gmm.last_log_etas = np.concatenate((gmm.last_log_etas, np.asarray([-1.0], dt)))
    gmm.num_received_updates = np.concatenate((gmm.num_received_updates, np.zeros(1, dt)))
    gmm.stepsizes = np.concatenate((gmm.stepsizes, np.asarray([gmm.initial_stepsize], dt)))
    gmm.reward_history = np.concatenate((gmm.reward_history, np.full((1, H), FLT_MIN, dt)), axis=0)
    gmm.weight_history = np.concatenate((gmm.weight_history, np.full((1, H), initial_weight, dt)), axis=0)


def remove_component(gmm: OracleGMM, idx: int):
    """GmmWrapper.remove_component (gmm_wrapper.py:129-148) + GMM.remove_component (gmm.py:388-399)."""
    keep = np.arange(gmm.num_components) != idx
    log_weights = gmm.log_weights[keep]
    gmm.means, gmm.chol_cov = gmm.means[keep], gmm.chol_cov[keep]
    gmm.replace_weights(log_weights)
    for name in ("l2_regularizers", "last_log_etas", "num_received_updates", "stepsizes", "reward_history",
                 "weight_history"):
        setattr(gmm, name, getattr(gmm, name)[keep])


class VipsComponentAdaptation:
    """component_adaptation.py:145-300 (without prior samples: num_prior_samples = 0, the default of every shipped
    configuration).  Randomness is injected: uniform_fn() -> the draw of tf.random.uniform([1]) (:208), shuffle_fn(n) ->
    the permutation of sample_db.get_random_sample (sample_db.py:151)."""

    def __init__(self, gmm: OracleGMM, db: OracleSampleDB, prior_mean, initial_cov, del_iters, add_iters, max_components,
                 thresholds_for_add_heuristic, min_weight_for_del_heuristic, num_database_samples):
        self.gmm, self.db = gmm, db
        dt, D = gmm.dt, gmm.num_dimensions
        # prior = DiagonalGMM(1, prior_mean, initial_cov): its average entropy (gmm.py:262-272, diagonal_gmm.py:36-37)
        self.prior_entropy = None
        if prior_mean is not None and initial_cov is not None:
            cov = np.broadcast_to(np.asarray(initial_cov, dt), (D,))
            self.prior_entropy = dt.type(0.5 * D * (LOG_2PI + 1)) + np.sum(np.log(np.sqrt(cov)))
        self.del_iters, self.add_iters, self.max_components = int(del_iters), int(add_iters), int(max_components)
        self.num_db_samples = num_database_samples
        self.num_calls_to_add_heuristic = 0
        self.thresholds = np.atleast_1d(np.asarray(thresholds_for_add_heuristic, dt))
        self.min_weight = min_weight_for_del_heuristic
        fd = int(math.floor(self.del_iters / 3))                                      # :171
        sigma = dt.type(self.del_iters / 8.0)
        xs = np.arange(-fd, fd).astype(dt)
        kernel = np.exp(-0.5 * (xs / sigma) ** 2) / (sigma * dt.type(math.sqrt(2 * math.pi)))      # Normal(0, sigma).prob
        self.kernel = kernel / np.sum(kernel)
        self.reward_improvements = np.zeros(0, dt)

    def adapt_number_of_components(self, iteration, uniform_fn, shuffle_fn, target):
        """:177-190.  Returns (deleted indices, index of the added component or None)."""
        deleted, added = [], None
        if iteration > self.del_iters:
            deleted = self.delete_bad_components()
        if iteration > 1 and iteration % self.add_iters == 0 and self.gmm.num_components < self.max_components:
            self.num_calls_to_add_heuristic += 1                                      # :242
            samples, lnpdfs = self.db.get_random_sample(self.num_db_samples, shuffle_fn)
            added = self.add_at_best_location(samples, lnpdfs, uniform_fn)
        return deleted, added

    def add_at_best_location(self, samples, target_lnpdfs, uniform_fn):
        """:193-226."""
        g = self.gmm
        dt, D = g.dt, g.num_dimensions
        it = self.num_calls_to_add_heuristic % len(self.thresholds)
        model_lq = log_density(g, samples)
        a = dt.type(uniform_fn())
        des_entropy = get_average_entropy(g) * a + self.prior_entropy * (1 - a) if self.prior_entropy is not None \
            else get_average_entropy(g)
        rewards = target_lnpdfs - np.maximum(np.max(model_lq) - self.thresholds[it], model_lq)
        new_mean = samples[int(np.argmax(rewards))]
        H_unscaled = dt.type(0.5 * D * (LOG_2PI + 1))
        c = np.exp((2 * (des_entropy - H_unscaled)) / D)
        new_cov = c * np.ones(D, dt) if g.diagonal_covs else c * np.eye(D, dtype=dt)
        add_component(g, 1e-29, new_mean, new_cov)
        return g.num_components - 1

    def delete_bad_components(self):
        """:261-300."""
        g = self.gmm
        ks, d = self.kernel.size, self.del_iters
        rh, wh = g.reward_history, g.weight_history
        current = np.mean(rh[:, -ks:] * self.kernel[None, :], axis=1)
        old = np.mean(rh[:, -ks - d:-d] * self.kernel[None, :], axis=1)
        old = old - np.max(current)
        current = current - np.max(current)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            self.reward_improvements = (current - old) / np.abs(old)
            max_actual = np.max(wh[:, -ks - d:-1], axis=1)
            window = rh[:, -ks - d:]
            max_greedy = np.max(np.exp(window - logsumexp(window, axis=0, keepdims=True)), axis=1)
        max_weights = np.maximum(max_actual, max_greedy)
        is_bad = (self.reward_improvements <= 0.4) & (max_weights < self.min_weight) & (rh[:, -d] != -FLT_MAX)
        bad = sorted(np.nonzero(is_bad)[0].tolist(), reverse=True)
        for idx in bad:
            remove_component(g, idx)
        return bad


# =====================================================================================================
# MMD evaluation (experiments/evaluation/mmd.py:4-78)
# =====================================================================================================
def mmd_sigma(groundtruth, max_points_for_median=1000, dtype=np.float32):
    """mmd.py:26-36 -> diagonal of the bandwidth matrix: per-dimension median of the squared differences of all pairs
    i <= j (i == j included) among the first `max_points_for_median` points.  tfp.stats.percentile(x, 50, axis=0)
    (tensorflow-probability 0.20.1, stats/quantiles.py, default interpolation 'nearest') sorts DESCENDING and takes
    element tf.round((d - 1) * (1 - q / 100)); tf.round rounds halves to even."""
    G = np.asarray(groundtruth, dtype)[:int(min(max_points_for_median, len(groundtruth)))]
    iu, ju = np.triu_indices(len(G))
    diff = (G[iu] - G[ju]) ** 2
    d = diff.shape[0]
    k_desc = int(np.round((d - 1) * 0.5))          # np.round: half to even
    return np.sort(diff, axis=0)[::-1][k_desc].astype(dtype)


def mmd_kernel_sum(X, Y, sigma_diag, alpha, dtype=np.float32):
    """mmd.py:38-56 (compute_ustat: Y = X; kernel_mix: X = ground truth): row by row like the reference,
    sum_i sum_j exp(-(x_i - y_j) K (x_i - y_j)) with K = inv(alpha * diag(sigma))."""
    X, Y = np.asarray(X, dtype), np.asarray(Y, dtype)
    kern = (1.0 / (dtype(alpha) * np.asarray(sigma_diag, dtype))).astype(dtype)
    total = dtype(0.0)
    for i in range(X.shape[0]):
        diff = X[i] - Y
        total = total + np.sum(np.exp(-np.sum(diff * kern * diff, axis=1)), dtype=dtype)
    return total


def mmd(groundtruth, model_sample, alpha, dtype=np.float32):
    """mmd.py:62-78."""
    sig = mmd_sigma(groundtruth, dtype=dtype)
    n1, n2 = len(groundtruth), len(model_sample)
    return (mmd_kernel_sum(groundtruth, groundtruth, sig, alpha, dtype) / n1 ** 2
            + mmd_kernel_sum(model_sample, model_sample, sig, alpha, dtype) / n2 ** 2
            - 2 * mmd_kernel_sum(groundtruth, model_sample, sig, alpha, dtype) / (n1 * n2))


__all__ = [n for n in dir() if not n.startswith("_")]
