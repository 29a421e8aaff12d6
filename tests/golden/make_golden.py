"""Generates tests/golden/samtron_small.npz: the oracle's trajectory on a small fixed SAMTRON problem.

The reference itself cannot be executed in this image (TensorFlow is not installable), so
the golden vectors come from the restated reference (oracle/, fp64 mode) with a seeded NumPy generator.  They pin
the oracle against accidental change and give the GPU tests a fixture that does not need the oracle's code path
to be re-derived.   Usage:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

K, D, DESIRED, ITERS = 4, 6, 50, 4


def problem():
    rng = np.random.default_rng(2024)
    means = rng.standard_normal((K, D)) * 2
    A = rng.standard_normal((K, D, D))
    covs = A @ A.transpose(0, 2, 1) / D + np.eye(D)
    tm = rng.standard_normal((3, D)) * 2
    tA = rng.standard_normal((3, D, D))
    tc = tA @ tA.transpose(0, 2, 1) / D + np.eye(D)
    noise = rng.standard_normal((ITERS, K * DESIRED, D))
    return means, covs, tm, tc, noise


def run_case():
    means, covs, tm, tc, noise = problem()
    g = O.make_full_gmm(np.ones(K) / K, means, covs, np.float64, initial_stepsize=0.1)
    target = O.gmm_target(np.ones(3) / 3, tm, tc, np.float64)
    db = O.OracleSampleDB(D, False, False, None, np.float64)
    cfg = O.IterationConfig(desired_samples_per_component=DESIRED, weight_stepsize=0.05)
    elbo = []
    out = {}
    for it in range(ITERS):
        E = noise[it]
        res = O.train_iter(g, db, target, cfg, lambda k, D_, n: E[k * DESIRED:(k + 1) * DESIRED].T)
        X, _ = O.sample_from_components_no_shuffle(g, [200] * K, lambda k, D_, n: noise[0][:200].T)
        lq = O.log_density(g, X)
        elbo.append(float(np.mean(target(X)[0] - lq)))
        if it == 0:
            out.update(bg0=res["bg"], H0=res["H_neg"], g0=res["g_neg"], elr0=res["elr"], etas0=res["update"]["etas"])
    out.update(means=g.means, chol=g.chol_cov, log_weights=g.log_weights, elbo=np.array(elbo),
               last_etas=g.last_log_etas)
    return out


if __name__ == "__main__":
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "samtron_small.npz"), **run_case())
    print("wrote samtron_small.npz")
