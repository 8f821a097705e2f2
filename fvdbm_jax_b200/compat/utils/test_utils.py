"""reference utils/test_utils.py: ``Key`` (a PRNG-key splitter, unused by every notebook) and ``print_dict``."""
import numpy as np

__all__ = ["Key", "print_dict"]


class Key:
    """Callable that hands out a fresh, reproducible generator per call (the reference splits a jax.random key)."""

    def __init__(self, seed):
        self._seq = np.random.SeedSequence(seed)

    def __call__(self):
        return np.random.default_rng(self._seq.spawn(1)[0])


def print_dict(dict: dict):  # noqa: A002  (the reference's parameter name)
    for k in dict:
        print(k + ": " + str(dict[k]))
