"""Running mean / variance (reference: Math/Statistics.py:8-40)."""
from __future__ import annotations


class OnlineEstimator:
    """Welford accumulator; the first sample is given to the constructor."""

    def __init__(self, x_):
        self.n = 1
        self.mean = x_ * 1.0
        self.m2 = x_ * 0.0

    def __call__(self, x_):
        self.n += 1
        delta = x_ - self.mean
        self.mean = self.mean + delta / self.n
        self.m2 = self.m2 + delta * (x_ - self.mean)
        return self.mean, self.m2 / (self.n - 1)
