"""Golden vectors of the two non-Gaussian TARGETS of the BASELINE configurations (C1 / C4: mixture of Student-t,
C2: planar robot), produced by the REFERENCE'S OWN classes (/root/reference/src/gmmvi/experiments/target_distributions/
student_t_mixture.py, planar_robot.py) imported unmodified over tests/golden/tf_shim (TensorFlow / TFP are not installable
here: the tfp.distributions they are written with are closed-form torch stand-ins, evaluated in float64; gradients come
from the reference's own GradientTape call, sample_selector.py:69-78).  Writes tests/golden/reference_targets.npz.

Usage (needs /root/reference; not run on the GPU box):  python tests/golden/make_reference_targets.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, "/root/reference/src")

import tensorflow as tf  # noqa: E402  (the shim)

tf.set_float_dtype(torch.float64)

from gmmvi.experiments.target_distributions.planar_robot import PlanarRobot  # noqa: E402
from gmmvi.experiments.target_distributions.student_t_mixture import StudentTMixture_LNPDF  # noqa: E402
from gmmvi.optimization.gmmvi_modules.sample_selector import SampleSelector  # noqa: E402


def f32(a):
    return np.asarray(a, np.float32).astype(np.float64)


def value_and_grad(target, X):
    """The reference's own way to get target gradients (SampleSelector.get_target_grads, sample_selector.py:69-78)."""
    sel = SampleSelector.__new__(SampleSelector)
    sel.target_distribution = target
    grad, val = sel.get_target_grads(tf.constant(X))
    return val.detach().numpy(), grad.detach().numpy()


def main():
    rng = np.random.default_rng(314)
    out = {}
    # ---- mixture of Student-t, constructed like make_target (student_t_mixture.py:154-169) -----------------------
    for name, D, ncomp, s in (("stm20", 20, 10, 20.0), ("stm_hard64", 64, 20, 25.0)):
        means = f32(rng.uniform(size=(ncomp, D)) * 2 * s - s)
        chols = np.empty((ncomp, D, D))
        for i in range(ncomp):
            a = 0.1 * D * rng.normal(0, 1, (D, D))
            cov = np.linalg.inv(a.T @ a + np.eye(D))
            chols[i] = f32(np.linalg.cholesky(cov))          # fp32-representable factors: covs = L L^T in float64
        covs = chols @ chols.transpose(0, 2, 1)
        w = np.ones(ncomp) / ncomp
        tgt = StudentTMixture_LNPDF(w, means, covs)
        # evaluation points: around the modes (tight, the covariances are ~1e-3) and far out in the heavy tails
        near = means[rng.integers(0, ncomp, 96)] + 0.05 * rng.standard_normal((96, D))
        mid = means[rng.integers(0, ncomp, 96)] + 2.0 * rng.standard_normal((96, D))
        far = s * 1.5 * rng.standard_normal((64, D))
        X = f32(np.concatenate([near, mid, far]))
        v, g = value_and_grad(tgt, X)
        out.update({f"{name}_means": means, f"{name}_chols": chols, f"{name}_X": X, f"{name}_lnpdf": v, f"{name}_grad": g,
                    f"{name}_marg3": tgt.marginal_log_density(tf.constant(X), 3).numpy()})
    # ---- planar robot (planar_robot.py:29-66; BASELINE C2 = 10 links, 4 goals) --------------------------------------
    for name, links, goals in (("planar10_4", 10, 4), ("planar10_1", 10, 1), ("planar3_4", 3, 4)):
        tgt = PlanarRobot(links, goals)
        X = f32(np.concatenate([rng.standard_normal((128, links)) * np.r_[1.0, 0.2 * np.ones(links - 1)],
                                rng.standard_normal((64, links))]))
        v, g = value_and_grad(tgt, X)
        out.update({f"{name}_X": X, f"{name}_lnpdf": v, f"{name}_grad": g,
                    f"{name}_fk": tgt.forward_kinematics(tf.constant(X)).numpy()})
    np.savez_compressed(os.path.join(HERE, "reference_targets.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
