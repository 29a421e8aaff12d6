"""CPU restatement of the OlegArenz/gmmvi hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.

PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path and
TensorFlow is not installable in this image, so the pin is built from the reference's own Python sources:
tests/golden/make_reference_golden.py executes /root/reference/src/gmmvi/{models,optimization}/... UNMODIFIED against
tests/golden/tf_shim (a torch-CPU stand-in for the ~75 TensorFlow / TFP ops the hot path uses, float64) on nine
configurations (+ the model surface) and commits every intermediate of every iteration (tests/golden/reference_*.npz); the restatement
reproduces all of them to <= 1e-13, mappings and sample counts bit exact (tests/test_oracle_pins.py::
test_oracle_matches_reference_sources).  What this cannot pin is TensorFlow's own kernels and random generators (replaced
by torch ops of the same documented semantics / by injected noise): in that sense the parity claim remains "the
reference's code over stand-in primitives", not "the reference's binaries".  The two non-Gaussian targets of the BASELINE configurations (mixture of Student-t, planar robot) are restated as well and
pinned the same way: tests/golden/make_reference_targets.py runs the reference's own StudentTMixture_LNPDF / PlanarRobot
classes over closed-form stand-ins for the tfp.distributions they use (tests/golden/reference_targets.npz), plus
scipy.stats.multivariate_t and finite differences.
Independent closed forms (scipy
``multivariate_normal.logpdf``, analytic Gaussian / categorical KL, Stein's identity on a Gaussian target, MORE on an
exactly quadratic target) are checked as well.
"""
from .gmmvi_oracle import *  # noqa: F401,F403
