"""Config assembly (mirror of configs/__init__.py:5-59).

The reference ships one YAML per module / experiment and merges them with `mergedeep`.  Here the same default
values live in `defaults.py` as plain dicts (the nested-dict schema IS the API: SURVEY.md section 5), and
`update_config` is a deep merge with mergedeep's Strategy.REPLACE semantics (nested dicts merge recursively,
leaves and lists are replaced)."""
from __future__ import annotations

import copy

import yaml

from .defaults import EXPERIMENT_CONFIGS, LETTER_TO_MODULE_CONFIG


def load_yaml(filename):
    """configs/__init__.py:5-11."""
    with open(filename, "r") as stream:
        try:
            return yaml.safe_load(stream)
        except yaml.YAMLError as exc:
            print(exc)


def _deep_merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _deep_merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def get_default_algorithm_config(algorithm_id):
    """configs/__init__.py:13-45: one module config per codeword letter, merged."""
    print(f"Using default parameters for codename {algorithm_id}")
    merged = dict()
    for letter in algorithm_id:
        _deep_merge(merged, copy.deepcopy(LETTER_TO_MODULE_CONFIG[letter.upper()]))
    return merged


def get_default_experiment_config(experiment_id):
    """configs/__init__.py:47-50."""
    print(f"Using default parameters for experiment {experiment_id}")
    if experiment_id not in EXPERIMENT_CONFIGS:
        raise FileNotFoundError(f"no default experiment config named '{experiment_id}'")
    return copy.deepcopy(EXPERIMENT_CONFIGS[experiment_id])


def get_default_config(algorithm_id, experiment_id):
    """configs/__init__.py:52-55."""
    return {**get_default_algorithm_config(algorithm_id), **get_default_experiment_config(experiment_id)}


def update_config(default_values, updates):
    """configs/__init__.py:57-59 (the top level is copied shallowly, exactly like the reference)."""
    return _deep_merge(dict(default_values), updates)
