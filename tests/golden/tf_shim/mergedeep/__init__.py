"""Stand-in for the two names of `mergedeep` that the reference's configs/__init__.py imports (test infrastructure, see
../tensorflow): recursive dict merge; with Strategy.REPLACE (also the default) non-dict values of the source replace the
destination's."""
from enum import Enum


class Strategy(Enum):
    REPLACE = 0


def merge(destination, *sources, strategy=Strategy.REPLACE):
    for src in sources:
        for k, v in src.items():
            if isinstance(v, dict) and isinstance(destination.get(k), dict):
                merge(destination[k], v, strategy=strategy)
            else:
                destination[k] = dict(v) if isinstance(v, dict) else v
    return destination
