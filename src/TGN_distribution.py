"""Truncated generalised normal sampler for pseudo-observation locations (reference: src/TGN_distribution.py:21-26).
The reference draws with adaptive rejection sampling (arspy, not installable here); the density is the same, sampled by
inversion of the truncated CDF.  Input-data generation only -- not on the accelerated path."""
import numpy as np
from scipy.special import gamma as Gamma
from scipy.stats import gennorm


def _scale(gamma, a, b):
    return Gamma(gamma) * abs(b - a) / 10.0


def log_TGN_pdf(x, gamma, alpha, a, b):
    s = _scale(gamma, a, b)
    mass = gennorm.cdf((b - alpha) / s, gamma) - gennorm.cdf((a - alpha) / s, gamma)
    return gennorm.logpdf((x - alpha) / s, gamma) - np.log(s * mass)


def TGN_sample(size, gamma, alpha, x_min, x_max):
    s = _scale(gamma, x_min, x_max)
    lo, hi = gennorm.cdf((x_min - alpha) / s, gamma), gennorm.cdf((x_max - alpha) / s, gamma)
    u = np.random.uniform(lo, hi, size)
    return alpha + s * gennorm.ppf(u, gamma)
