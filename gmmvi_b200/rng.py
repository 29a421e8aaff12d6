"""Process-wide seed / subsequence bookkeeping for the counter-based device generator.

The reference seeds TF/NumPy/Python with tf.keras.utils.set_random_seed (gmmvi_runner.py:38).  Here every
draw of standard-normal noise is `gvi_fill_normal_f32(seed, subsequence, global_row)`; `subsequence`
increases by one per draw so that successive draws are independent, and the value of a sample depends
only on its GLOBAL row index -- the property multi-GPU sharding relies on.
"""
from __future__ import annotations

import random

import numpy as np
import torch

_state = {"seed": 0, "subsequence": 0}


def set_seed(seed: int):
    seed = int(seed)
    _state["seed"] = seed
    _state["subsequence"] = 0
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)


def seed() -> int:
    return _state["seed"]


def next_subsequence() -> int:
    s = _state["subsequence"]
    _state["subsequence"] = s + 1
    return s
