"""Restatement of the reference DR samplers and their exact target laws.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restated from /root/reference/random_envs/random_env.py:
  uniform     :150-151   lo + (hi-lo)*U[0,1) per dim, independent
  truncnorm   :153-171   X = mean + std*TN(-2,2) (scipy ``truncnorm.rvs`` == inverse-CDF
                         of one uniform); while X < lb_i redraw; after the 3rd redraw the
                         value is replaced by lb_i (the 4th draw is consumed, then discarded)
  gaussian    :173-190   X = randn*std + mean; while X < 0.1 (hard-coded) redraw; if the
                         first three draws are all < 0.1 a 4th is consumed and then
                         ``Exception('Not all samples were above > 0.1 after 2 attempts')``
  interleave  :102-121   distr = [a0, b0, a1, b1, ...]

RNG streams cannot be matched with a counter-based GPU generator (the reference
uses the process-global numpy/scipy state), so what parity means for samplers is
(i) the retry/clip control flow is identical for identical draws -- tested with a
scripted draw source against the real reference -- and (ii) the output LAW is
identical: the CDFs below are the exact laws implied by the code above and are
what ``scipy.stats.kstest`` is run against for the GPU samples.
"""
import math

import numpy as np
from scipy import special, stats

GAUSSIAN_FLOOR = 0.1           # random_env.py:181 (hard-coded, not get_task_lower_bound)
TN_A, TN_B = -2.0, 2.0         # random_env.py:154
GAUSSIAN_ERROR = "Not all samples were above > 0.1 after 2 attempts"   # random_env.py:186

# closed-form moments of the standard normal truncated to [-2, 2]
_Z = special.ndtr(TN_B) - special.ndtr(TN_A)
_PHI2 = math.exp(-2.0) / math.sqrt(2.0 * math.pi)
TN_VAR = 1.0 - 2.0 * TN_B * _PHI2 / _Z                   # 0.7737413035499232
TN_M4 = 3.0 - 2.0 * (TN_B ** 3 + 3.0 * TN_B) * _PHI2 / _Z    # E[X^4] of the standard TN(-2,2)
TN_EXCESS_KURT = TN_M4 / TN_VAR ** 2 - 3.0                # -0.6344632828703505


def split_interleaved(distr):
    """random_env.py:102-121: [a0,b0,a1,b1,...] -> (a, b); loop is range(len//2)."""
    k = len(distr) // 2
    return np.array([distr[2 * i] for i in range(k)], float), np.array([distr[2 * i + 1] for i in range(k)], float)


def sample_uniform(lo, hi, u):
    """u: array of U[0,1) draws, shape (..., dim).  numpy's ``uniform`` is lo + (hi-lo)*u."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    return lo + (hi - lo) * u


def sample_truncnorm_dim(mean, std, lb, draw):
    """One dim of random_env.py:156-169.  ``draw()`` returns one TN(-2,2,loc,scale) variate."""
    attempts = 0
    obs = draw()
    while obs < lb:
        obs = draw()
        attempts += 1
        if attempts > 2:
            obs = lb
    return obs


def sample_gaussian_dim(mean, std, draw):
    """One dim of random_env.py:177-188.  ``draw()`` returns one standard normal."""
    attempts = 0
    obs = draw() * std + mean
    while obs < GAUSSIAN_FLOOR:
        obs = draw() * std + mean
        attempts += 1
        if attempts > 2:
            raise Exception(GAUSSIAN_ERROR)
    return obs


def sample_task(dr_type, a, b, lb=None, rng=None):
    """Whole-vector port of ``RandomEnv.sample_task`` on an explicit RandomState."""
    rng = rng or np.random
    a, b = np.asarray(a, float), np.asarray(b, float)
    if dr_type == "uniform":
        return rng.uniform(a, b, a.shape)
    if dr_type == "truncnorm":
        return np.array([sample_truncnorm_dim(m, s, l, lambda m=m, s=s: truncnorm_ppf(rng.uniform(), m, s))
                         for m, s, l in zip(a, b, lb)])
    if dr_type == "gaussian":
        return np.array([sample_gaussian_dim(m, s, rng.randn) for m, s in zip(a, b)])
    raise ValueError("sampling value of random env needs to be set before using sample_task() "
                     "or set_random_task(). Set it by uploading a DR distr.")


def truncnorm_ppf(u, mean, std):
    """Inverse CDF of mean + std*TN(-2,2) -- what ``truncnorm.rvs`` evaluates on one uniform."""
    pa = special.ndtr(TN_A)
    return mean + std * special.ndtri(pa + u * (special.ndtr(TN_B) - pa))


# ----------------------------------------------------------------------------- target laws
def uniform_cdf(lo, hi):
    return lambda x: np.clip((np.asarray(x, float) - lo) / (hi - lo), 0.0, 1.0)


def truncnorm_lb_cdf(mean, std, lb):
    """Law of one truncnorm dim INCLUDING the lower-bound mixture.

    With F the CDF of mean+std*TN(-2,2) and p = F(lb): the first of three draws that is
    >= lb is kept, else lb.  P(X <= x) = p^3 + (1+p+p^2) (F(x) - p) for x >= lb, 0 below.
    """
    base = stats.truncnorm(TN_A, TN_B, loc=mean, scale=std)
    p = float(base.cdf(lb))

    def cdf(x):
        x = np.asarray(x, float)
        f = base.cdf(x)
        return np.where(x < lb, 0.0, p ** 3 + (1.0 + p + p * p) * (f - p))
    return cdf, p


def gaussian_floor_cdf(mean, std):
    """Law of one gaussian dim conditional on no exception: normal truncated below at 0.1.

    Draws are iid and the first of three that is >= 0.1 is kept, so conditional on success
    the law is N(mean, std) restricted to [0.1, inf).  P(exception) = q^3, q = Phi((0.1-mean)/std).
    """
    q = float(stats.norm.cdf(GAUSSIAN_FLOOR, mean, std))

    def cdf(x):
        x = np.asarray(x, float)
        return np.where(x < GAUSSIAN_FLOOR, 0.0, (stats.norm.cdf(x, mean, std) - q) / (1.0 - q))
    return cdf, q


def ks_distance(samples, cdf):
    """sup_x |F_n(x) - F(x)| that stays valid when F has atoms (the truncnorm lower-bound mass).

    ``scipy.stats.kstest`` assumes a continuous F and reports the atom's jump as the distance.
    Here both the right limits (F_n(v), F(v)) and the left limits (F_n(v-), F(v-)) are compared at
    every distinct sample value v, F(v-) being evaluated one ulp below v.
    """
    x = np.sort(np.asarray(samples, dtype=np.float64))
    n = x.size
    v, counts = np.unique(x, return_counts=True)
    right = np.cumsum(counts) / n
    left = right - counts / n
    f_right = cdf(v)
    f_left = cdf(np.nextafter(v, -np.inf))
    return float(max(np.max(np.abs(right - f_right)), np.max(np.abs(left - f_left))))


# ----------------------------------------------------------------------------- fullgaussian (random_env.py:192-220)
def denormalize(p, lo, hi):
    """``denormalize_parameters`` (:205-220): (p * (hi - lo)) / 4 + lo, p in the normalised [0, 4] space."""
    p, lo, hi = np.asarray(p, float), np.asarray(lo, float), np.asarray(hi, float)
    return (p * (hi - lo)) / 4 + lo


def sample_fullgaussian_from_z(mean, factor, lo, hi, z):
    """The law of ``sample_task`` for 'fullgaussian' on explicit standard normals z (..., dim):
    multivariate normal in the normalised space (:194), clip to [0, 4] (:195), denormalise (:197)."""
    x = np.asarray(mean, float) + np.asarray(z, float) @ np.asarray(factor, float).T
    return denormalize(np.clip(x, 0, 4), lo, hi)


def box_muller_f64(u_radius_open0, u_angle):
    """The framework's fp64 normal pair: sqrt(-2 ln u1) * (cos, sin)(2 pi u2), u1 in (0,1], u2 in [0,1)."""
    rad = np.sqrt(-2.0 * np.log(u_radius_open0))
    return rad * np.cos(2.0 * np.pi * u_angle), rad * np.sin(2.0 * np.pi * u_angle)
