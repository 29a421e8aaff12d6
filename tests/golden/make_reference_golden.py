"""Golden vectors produced by the REFERENCE'S OWN PYTHON SOURCES (/root/reference/src/gmmvi), executed unmodified in
this container against tests/golden/tf_shim (a torch-CPU stand-in for the TensorFlow ops the hot path uses: TensorFlow
itself is not installable here).  Writes tests/golden/reference_<case>.npz; tests/test_oracle_pins.py
checks the oracle (oracle/gmmvi_oracle.py) against them, the GPU tests check the device path against the oracle.

What this pins: the reference's control flow, formulas and quirks (module order, double normalisation of the importance
weights, bracketing searches and their stop rules, l2 / eta bookkeeping, first-occurrence unique, ...), evaluated in
float64 so that rounding does not blur the comparison.  What it cannot pin: TensorFlow's own kernels (replaced by torch
ops of the same documented semantics) and its random generators (noise is injected from a seeded NumPy generator).

Usage (needs /root/reference; not run on the GPU box):  python tests/golden/make_reference_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, "/root/reference/src")
if ROOT not in sys.path:
    sys.path.append(ROOT)
np.bool = bool          # least_squares.py:111 uses np.bool, removed from NumPy >= 1.24

import tensorflow as tf  # noqa: E402  (the shim)

tf.set_float_dtype(torch.float64)

from gmmvi.experiments.target_distributions.lnpdf import LNPDF  # noqa: E402
from gmmvi.models.diagonal_gmm import DiagonalGMM  # noqa: E402
from gmmvi.models.full_cov_gmm import FullCovGMM  # noqa: E402
from gmmvi.models.gmm_wrapper import GmmWrapper  # noqa: E402
from gmmvi.optimization.gmmvi import GMMVI  # noqa: E402



class GmmTarget(LNPDF):
    """A Gaussian-mixture target written against the reference's LNPDF interface with the reference's own FullCovGMM
    (what experiments/target_distributions/gmm.py does with tfp distributions)."""

    def __init__(self, weights, means, covs):
        super().__init__(use_log_density_and_grad=False, safe_for_tf_graph=True)
        self.gmm = FullCovGMM(tf.constant(weights), tf.constant(means), tf.constant(covs))

    def get_num_dimensions(self):
        return int(self.gmm.num_dimensions)

    def log_density(self, x):
        return self.gmm.log_density(x)


from cases import CASES, DESIRED, K, D, base_config  # noqa: E402,F401  (pure data, shared with the tests)


def f32(a):
    """Round to float32 and return as float64: every input is exactly representable on the device."""
    return np.asarray(a, np.float32).astype(np.float64)


def problem(diagonal, K, D, seed=2025):
    """Initial mixture and target with fp32-representable parameters.  Covariances are L L^T of fp32-representable
    factors, so the reference's (float64) Cholesky returns L to 1e-16 and the device can be started from exactly the same
    factor (FullCovGMM.from_cholesky)."""
    rng = np.random.default_rng(seed)
    means = f32(rng.standard_normal((K, D)) * 2)
    A = rng.standard_normal((K, D, D))
    chols = f32(np.linalg.cholesky(A @ A.transpose(0, 2, 1) / D + np.eye(D)))
    if diagonal:
        chols = np.stack([np.diag(np.diag(c)) for c in chols])
    tm = f32(rng.standard_normal((3, D)) * 2)
    tA = rng.standard_normal((3, D, D))
    tchols = f32(np.linalg.cholesky(tA @ tA.transpose(0, 2, 1) / D + np.eye(D)))
    return means, chols, tm, tchols


def run_reference(name):
    over, iters, diagonal = CASES[name][:3]
    K_, D_ = CASES[name][3] if len(CASES[name]) > 3 else (K, D)
    cfg = base_config(**over)
    means, chols, tm, tchols = problem(diagonal, K_, D_)
    covs, tc = chols @ chols.transpose(0, 2, 1), tchols @ tchols.transpose(0, 2, 1)
    w = np.ones(K_) / K_
    if diagonal:
        model = DiagonalGMM(tf.constant(w), tf.constant(means), tf.constant(np.stack([np.diag(c) for c in covs])))
    else:
        model = FullCovGMM(tf.constant(w), tf.constant(means), tf.constant(covs))
    target = GmmTarget(np.ones(3) / 3, tm, tc)
    wrapped = GmmWrapper.build_from_config(model, cfg)
    gmmvi = GMMVI.build_from_config(cfg, target, wrapped)
    rng = np.random.default_rng(7)
    draws = []                                   # every (D, n) block of standard-normal noise, in call order

    def normal_hook(shape):
        e = f32(rng.standard_normal(shape))
        draws.append(e)
        return e
    tf.random.normal_hook = normal_hook
    uniforms, perms = [], []                     # draws of the component adaptation, in call order

    def uniform_hook(shape):
        u = f32(rng.uniform(size=shape))
        uniforms.append(u.reshape(-1))
        return u

    def shuffle_hook(n):
        p = rng.permutation(n)
        perms.append(p)
        return p
    tf.random.uniform_hook, tf.random.shuffle_hook = uniform_hook, shuffle_hook
    out = {"init_means": means, "init_chols": chols, "init_covs": covs, "target_means": tm, "target_chols": tchols,
           "target_covs": tc}
    big = D_ > 32          # keep the fixture small: samples / gradients are derivable and are checked in the small cases
    for it in range(iters):
        n0, u0, p0 = len(draws), len(uniforms), len(perms)
        samples, mapping, bg, lnpdfs, grads = gmmvi.sample_selector.select_samples()
        H, g = gmmvi.ng_estimator.get_expected_hessian_and_grad(samples, mapping, bg, lnpdfs, grads)
        gmmvi._run_updates(samples, mapping, bg, lnpdfs, grads)
        gmmvi.num_component_adapter.adapt_number_of_components(gmmvi.num_updates)
        m = gmmvi.model
        out.update({
            f"noise_shapes{it}": np.array([d.shape for d in draws[n0:]], dtype=np.int64).reshape(-1, 2),
            f"noise{it}": np.concatenate([d.T for d in draws[n0:]], axis=0) if len(draws) > n0 else np.zeros((0, D_)),
            f"samples{it}": samples.numpy(), f"mapping{it}": mapping.numpy().astype(np.int64), f"bg{it}": bg.numpy(),
            f"lnpdfs{it}": lnpdfs.numpy(), f"grads{it}": grads.numpy(), f"H{it}": H.numpy(), f"g{it}": g.numpy(),
            f"means{it}": m.means.numpy().copy(), f"chol{it}": m.chol_cov.numpy().copy(),
            f"log_weights{it}": m.log_weights.numpy().copy(), f"stepsizes{it}": m.stepsizes.numpy().copy(),
            f"l2{it}": m.l2_regularizers.numpy().copy(), f"last_log_etas{it}": m.last_log_etas.numpy().copy(),
            f"num_received_updates{it}": m.num_received_updates.numpy().copy(),
        })
        if cfg["sample_selector_type"] == "mixture-based":
            out[f"uniform{it}"] = np.concatenate(uniforms[u0:]) if len(uniforms) > u0 else np.zeros(0)
        if cfg["num_component_adapter_type"] == "adaptive":
            out.update({f"uniform{it}": np.concatenate(uniforms[u0:]) if len(uniforms) > u0 else np.zeros(0),
                        f"perm{it}": np.concatenate(perms[p0:]).astype(np.int32) if len(perms) > p0 else np.zeros(0, np.int32),
                        f"reward_history{it}": m.reward_history.numpy().copy(),
                        f"weight_history{it}": m.weight_history.numpy().copy()})
        if cfg["num_component_adapter_type"] == "adaptive":
            for key in ("samples", "grads"):            # derivable; checked in the fixed-K cases
                del out[f"{key}{it}"]
            out[f"noise{it}"] = out[f"noise{it}"].astype(np.float32)
            out[f"mapping{it}"] = out[f"mapping{it}"].astype(np.int32)
        if big:
            for key in ("samples", "grads"):
                del out[f"{key}{it}"]
            out.pop("init_covs", None)              # = chols chols^T
            out.pop("target_covs", None)
            out[f"noise{it}"] = out[f"noise{it}"].astype(np.float32)        # exactly representable (f32() above)
    out["iterations"] = np.array(iters)
    return out


def run_model_api(diagonal):
    """Direct calls of the model surface of the reference's FullCovGMM / DiagonalGMM."""
    K_, D_, N_ = 5, 7, 40
    means, chols, _, _ = problem(diagonal, K_, D_, seed=77)
    rng = np.random.default_rng(78)
    w = f32(rng.dirichlet(np.ones(K_)))
    X = f32(rng.standard_normal((N_, D_)) * 2.5)
    covs = chols @ chols.transpose(0, 2, 1)
    if diagonal:
        model = DiagonalGMM(tf.constant(w), tf.constant(means), tf.constant(np.stack([np.diag(c) for c in covs])))
    else:
        model = FullCovGMM(tf.constant(w), tf.constant(means), tf.constant(covs))
    Xt = tf.constant(X)
    out = {"init_means": means, "init_chols": chols, "weights_in": w, "X": X}
    out["log_weights"] = model.log_weights.numpy().copy()
    out["component_log_densities"] = model.component_log_densities(Xt).numpy()
    lq, lqk = model.log_densities_also_individual(Xt)
    out["log_density"], out["individual"] = lq.numpy(), lqk.numpy()
    out["log_density_only"] = model.log_density(Xt).numpy()
    out["density"] = model.density(Xt).numpy()
    lq2, grad, lqk2 = model.log_density_and_grad(tf.constant(X))
    out["lq_grad"], out["grad"], out["lqk_grad"] = lq2.numpy(), grad.numpy(), lqk2.numpy()
    out["component_entropies"] = model.component_entropies().numpy()
    out["average_entropy"] = np.asarray(model.get_average_entropy().numpy())
    out["covs"] = model.covs.numpy()
    if not diagonal:
        out["component_log_density_2"] = model.component_log_density(2, Xt).numpy()
        cl, cg = model.component_log_density_and_grad(1, tf.constant(X))
        out["component_lq_1"], out["component_grad_1"] = cl.numpy(), cg.numpy()
        out["component_marginal_3"] = model.component_marginal_log_densities(Xt, 3).numpy()
        out["marginal_3"] = model.marginal_log_density(Xt, 3).numpy()
    # sampling: categorical, GMM.sample (quirk 4: samples grouped by component, component indices in draw order),
    # per-component counts.  (Quirk 5 - an all-False comparison row selects component 0 - needs u >= cumsum(w)[-1],
    # which depends on the last ulp of the cumulative sum: it is tested on the oracle and the device directly.)
    u = f32(rng.uniform(size=(30, 1)))
    u[0, 0] = np.nextafter(np.float32(1.0), np.float32(0.0))
    tf.random.uniform_hook = lambda shape: u
    out["u"] = u[:, 0]
    out["sample_categorical"] = model.sample_categorical(30).numpy().astype(np.int32)
    draws = []

    def normal_hook(shape):
        e = f32(rng.standard_normal(shape))
        draws.append(e)
        return e
    tf.random.normal_hook = normal_hook
    xs, comps = model.sample(30)
    out["sample_x"], out["sample_components"] = xs.numpy(), comps.numpy().astype(np.int32)
    out["sample_noise"] = np.concatenate([d.T for d in draws], axis=0)
    out["sample_noise_shapes"] = np.array([d.shape for d in draws], dtype=np.int64)
    draws.clear()
    n_per = np.array([3, 0, 5, 1, 2], dtype=np.int64)
    xs2, mapping = model.sample_from_components_no_shuffle(tf.constant(n_per))
    out["n_per"], out["no_shuffle_x"], out["no_shuffle_mapping"] = n_per, xs2.numpy(), mapping.numpy().astype(np.int32)
    out["no_shuffle_noise"] = np.concatenate([d.T for d in draws], axis=0)
    # structural edits: replace_weights (normalises), add_component, remove_component
    model.replace_weights(tf.constant(f32(np.log(w) + rng.standard_normal(K_))))
    out["log_weights_replaced"] = model.log_weights.numpy().copy()
    new_cov = f32(np.diag(covs[0]) * 0.5) if diagonal else f32(covs[0] * 0.5 + np.eye(D_))
    model.add_component(tf.constant(f32(0.2)), tf.constant(f32(means[0] + 1.0)), tf.constant(new_cov))
    out["new_cov"] = new_cov
    out["added_log_weights"], out["added_chol"] = model.log_weights.numpy().copy(), model.chol_cov.numpy().copy()
    model.remove_component(1)
    out["removed_log_weights"], out["removed_means"] = model.log_weights.numpy().copy(), model.means.numpy().copy()
    return out


if __name__ == "__main__":
    only = sys.argv[1:] or list(CASES) + ["model_api_full", "model_api_diagonal"]
    for name in only:
        res = run_model_api(name.endswith("diagonal")) if name.startswith("model_api") else run_reference(name)
        np.savez_compressed(os.path.join(HERE, f"reference_{name}.npz"), **res)
        print(f"wrote reference_{name}.npz  ({len(res)} arrays)")
