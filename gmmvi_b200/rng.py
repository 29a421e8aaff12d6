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
# While a CUDA graph of the iteration is captured (optimization/graphed.py) the draw counter lives on the device: a draw
# uses *counter + (draws so far in this capture), and the graph ends with counter += draws, so every replay sees fresh
# subsequences -- the same ones an eager run would have used.
_device = {"counter": None, "draws": 0}


def begin_device_mode(counter):
    """counter: int64 device tensor [1] holding the subsequence of the next draw."""
    _device["counter"], _device["draws"] = counter, 0


def end_device_mode() -> int:
    """-> number of draws issued since begin_device_mode."""
    n = _device["draws"]
    _device["counter"], _device["draws"] = None, 0
    return n


def device_counter():
    """(counter tensor, offset) for the next draw in device mode, or None in host mode; advances the offset."""
    if _device["counter"] is None:
        return None
    off = _device["draws"]
    _device["draws"] = off + 1
    return _device["counter"], off


def advance(n: int):
    """Account on the host for `n` draws made by a graph replay."""
    _state["subsequence"] += int(n)


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
