"""CPU restatement of the OlegArenz/gmmvi hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
path and TensorFlow is not installable in this image, so the
restatement is pinned only against independent closed forms (scipy
``multivariate_normal.logpdf``, analytic Gaussian / categorical KL, Stein's identity on
a Gaussian target, MORE on an exactly quadratic target) -- see tests/test_oracle_pins.py.
"""
from .gmmvi_oracle import *  # noqa: F401,F403
