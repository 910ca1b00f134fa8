"""T3: DR sampler kernels -- exact draws vs the oracle's Philox contract, moments + KS vs the target laws.

KS critical value at alpha = 0.001 is 1.95 / sqrt(n) (SURVEY.md section 8d cfg 5).  Moment tolerances are
6 standard errors of the respective estimator.
"""
import os

import numpy as np
import pytest
import torch
from scipy import stats

import random_envs_b200 as random_envs
from oracle import c_oracle, dr_port

pytestmark = pytest.mark.gpu

NU = np.array(random_envs.HUMANOID_NOMINAL)
N_BIG = 1 << 20
KS_SUB = 1 << 18


def _interleave(a, b):
    return list(np.stack([a, b], 1).reshape(-1))


def _sampler(env_id, dr_type, a, b, seed=0):
    s = random_envs.TaskSampler(env_id)
    s.seed_dr(seed)
    s.set_dr_distribution(dr_type, _interleave(a, b))
    return s


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("env_id", ["RandomHopperUnmodeled-v0", "RandomCartPole-v0", "RandomWalker2d-v0", "RandomHumanoid-v0"])
def test_uniform_draws_are_bit_exact_vs_oracle_contract(env_id, dtype):
    table = random_envs.XI_TABLES[env_id]
    lo = np.array([b[0] for b in table.search_bounds]); hi = np.array([b[1] for b in table.search_bounds])
    s = _sampler(env_id, "uniform", lo, hi, seed=77)
    x = s.sample_tasks_tensor(1000, dtype=dtype).cpu().numpy()            # call index 0
    y = s.sample_tasks_tensor(300, dtype=dtype).cpu().numpy()             # call index 1
    npdt = np.float32 if dtype == torch.float32 else np.float64
    for i in (0, 1, 127, 128, 999):
        assert np.array_equal(x[i], c_oracle.xi_uniform(77, i, 0, lo, hi, purpose=c_oracle.PURPOSE_TASKS, dtype=npdt)), i
    assert np.array_equal(y[5], c_oracle.xi_uniform(77, 5, 1, lo, hi, purpose=c_oracle.PURPOSE_TASKS, dtype=npdt))
    assert x.shape == (1000, len(lo)) and np.all(x >= lo) and np.all(x <= hi)
    # launch-geometry invariance: a longer call reproduces the shorter one row for row
    s2 = _sampler(env_id, "uniform", lo, hi, seed=77)
    assert np.array_equal(s2.sample_tasks_tensor(4097, dtype=dtype).cpu().numpy()[:1000], x)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_truncnorm_transform_matches_scipy_on_the_oracles_uniforms(dtype):
    """Same uniforms (oracle Philox) pushed through scipy's ppf must give the kernel's values."""
    mean, std = NU, 0.1 * NU
    s = _sampler("RandomHumanoid-v0", "truncnorm", mean, std, seed=9)
    x = s.sample_tasks_tensor(64, dtype=dtype).cpu().numpy().astype(np.float64)
    npdt = np.float32 if dtype == torch.float32 else np.float64
    for i in (0, 17, 63):
        u = c_oracle.uniforms(9, i, 0, c_oracle.PURPOSE_TASKS, 0, 30, dtype=npdt).astype(np.float64)
        want = dr_port.truncnorm_ppf(u, mean, std)
        tol = 2e-5 if dtype == torch.float32 else 1e-12
        assert np.max(np.abs(x[i] - want) / std) <= tol, (i, np.max(np.abs(x[i] - want) / std))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dr_type", ["uniform", "gaussian", "truncnorm"])
def test_humanoid_30dim_moments_and_ks(dr_type, dtype):
    """BASELINE config 5 at test scale: 2^20 x 30 samples per call, moments vs closed form, KS per dim."""
    n = N_BIG
    if dr_type == "uniform":
        a, b = 0.5 * NU, 1.5 * NU
    else:
        a, b = NU, 0.1 * NU
    s = _sampler("RandomHumanoid-v0", dr_type, a, b, seed=3)
    x = s.sample_tasks_tensor(n, dtype=dtype)
    s.check_dr_violations()
    assert x.shape == (n, 30) and x.is_cuda and x.dtype == dtype
    xd = x.double()
    mean, var = xd.mean(0).cpu().numpy(), xd.var(0).cpu().numpy()
    z = (xd - xd.mean(0)) / xd.std(0)
    kurt = (z ** 4).mean(0).cpu().numpy() - 3.0
    skew = (z ** 3).mean(0).cpu().numpy()
    if dr_type == "uniform":
        m0, v0, k0 = (a + b) / 2, (b - a) ** 2 / 12, -1.2
    elif dr_type == "gaussian":
        m0, v0, k0 = a, b ** 2, 0.0           # floor 0.1 is >= 8.5 sigma away for every dim: no truncation
    else:
        m0, v0, k0 = a, dr_port.TN_VAR * b ** 2, dr_port.TN_EXCESS_KURT
    se_mean = np.sqrt(v0 / n)
    assert np.all(np.abs(mean - m0) <= 6 * se_mean + 1e-6 * np.abs(m0)), np.max(np.abs(mean - m0) / se_mean)
    assert np.all(np.abs(var / v0 - 1.0) <= 6 * np.sqrt((k0 + 2.0) / n) + 1e-5)
    assert np.all(np.abs(skew) <= 6 * np.sqrt(6.0 / n) * 2) and np.all(np.abs(kurt - k0) <= 6 * np.sqrt(24.0 / n) * 2)
    sub = x[:KS_SUB].cpu().numpy().astype(np.float64)
    crit = 1.95 / np.sqrt(KS_SUB)
    lb = random_envs.XI_TABLES["RandomHumanoid-v0"].lower_bounds
    for d in range(30):
        if dr_type == "uniform":
            cdf = dr_port.uniform_cdf(a[d], b[d])
        elif dr_type == "gaussian":
            cdf = dr_port.gaussian_floor_cdf(a[d], b[d])[0]
        else:
            cdf = dr_port.truncnorm_lb_cdf(a[d], b[d], lb[d])[0]
        assert dr_port.ks_distance(sub[:, d], cdf) < crit, (dr_type, d)
    if dr_type == "truncnorm":
        assert np.all(sub >= a - 2 * b - 1e-6 * a) and np.all(sub <= a + 2 * b + 1e-6 * a)
    # independence across dims (different Philox lanes/blocks): sample correlations ~ N(0, 1/n)
    c = np.corrcoef(sub[:, :8].T)
    assert np.max(np.abs(c - np.eye(8))) < 6 / np.sqrt(KS_SUB)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_truncnorm_lower_bound_point_mass_and_gaussian_floor(dtype, golden_dir):
    """The cartpole case of the golden reference draws: pole_mass mean == lb -> mass 1/8 exactly at lb."""
    g = np.load(os.path.join(golden_dir, "sampler_reference_draws.npz"))
    n = 1 << 19
    s = random_envs.TaskSampler("RandomCartPole-v0"); s.seed_dr(21)
    s.set_dr_distribution("truncnorm", list(g["truncnorm_params"]))
    x = s.sample_tasks_tensor(n, dtype=dtype).cpu().numpy().astype(np.float64)
    mu, sd = dr_port.split_interleaved(list(g["truncnorm_params"]))
    crit = 1.95 / np.sqrt(n)
    lb32 = float(np.float32(0.1)) if dtype == torch.float32 else 0.1
    for d in range(4):
        cdf, p = dr_port.truncnorm_lb_cdf(mu[d], sd[d], 0.1)
        xs = np.where(x[:, d] == lb32, 0.1, x[:, d])        # the fp32 image of the bound is the atom
        assert dr_port.ks_distance(xs, cdf) < crit + 2e-7 / sd[d], d
        assert xs.min() >= 0.1
        mass = np.mean(xs == 0.1)
        assert abs(mass - p ** 3) <= 5 * np.sqrt(max(p ** 3 * (1 - p ** 3), 1e-12) / n) + 1e-9, (d, mass, p ** 3)
    assert abs(np.mean(x[:, 2] == lb32) - 0.125) < 0.003
    # two-sample KS against the reference's OWN draws (4000 iid samples of its sample_task)
    xa = np.where(x == lb32, 0.1, x)                        # fp32: map the atom back onto the fp64 bound
    for d in range(4):
        assert stats.ks_2samp(xa[:20000, d], g["truncnorm"][:, d]).pvalue > 1e-4, d
    # gaussian with pole_mass floor 2 sigma below the mean: conditional law, no exception expected to be likely
    s.set_dr_distribution("gaussian", list(g["gaussian_params"]))
    y = s.sample_tasks_tensor(n, dtype=dtype).cpu().numpy().astype(np.float64)
    mu, sd = dr_port.split_interleaved(list(g["gaussian_params"]))
    for d in range(4):
        cdf, q = dr_port.gaussian_floor_cdf(mu[d], sd[d])
        assert dr_port.ks_distance(y[:, d], cdf) < crit + 2e-7 / sd[d], d
        assert stats.ks_2samp(y[:20000, d], g["gaussian"][:, d]).pvalue > 1e-4, d
    assert y.min() >= lb32 - 1e-12


def test_gaussian_three_failures_raise_the_reference_exception():
    s = random_envs.TaskSampler("RandomCartPole-v0")
    s.set_dr_distribution("gaussian", [9.8, 1.0, 1.0, 0.1, -5.0, 0.1, 0.5, 0.05])     # pole_mass can never reach 0.1
    with pytest.raises(Exception, match="Not all samples were above > 0.1 after 2 attempts"):
        s.sample_task()
    s.set_dr_distribution("gaussian", [9.8, 1.0, 1.0, 0.1, 0.2, 0.01, 0.5, 0.05])
    x = s.sample_task()                                                               # counter was cleared
    assert x.shape == (4,) and x.dtype == np.float64 and np.all(x > 0.1)


def test_numpy_api_shapes_and_reference_snippet():
    env = random_envs.gym.make("RandomCartPole-v0")
    env.set_dr_distribution(dr_type="uniform", distr=[2, 20, 0.5, 3, 0.05, 0.3, 0.1, 1.0])
    x = env.sample_task()
    assert isinstance(x, np.ndarray) and x.shape == (4,) and x.dtype == np.float64
    xs = env.sample_tasks(7)
    assert xs.shape == (7, 4) and not np.array_equal(xs[0], x)        # a new call index, a new stream
    env.set_random_task()
    t = env.get_task()
    assert np.all(t >= [2, 0.5, 0.05, 0.1]) and np.all(t <= [20, 3, 0.3, 1.0]) and env.total_mass == t[1] + t[2]


@pytest.mark.parametrize("env_id", sorted(random_envs.XI_TABLES))
def test_every_env_of_the_suite_samples_inside_its_search_bounds(env_id):
    table = random_envs.XI_TABLES[env_id]
    lo = np.array([b[0] for b in table.search_bounds]); hi = np.array([b[1] for b in table.search_bounds])
    s = _sampler(env_id, "uniform", lo, hi, seed=1)
    x = s.sample_tasks(513)
    assert x.shape == (513, len(lo)) and np.all(x >= lo) and np.all(x < hi)
    mid = (lo + hi) / 2
    s.set_dr_distribution("truncnorm", _interleave(mid, (hi - lo) / 8))
    y = s.sample_tasks(513)
    assert np.all(np.abs(y - mid) <= (hi - lo) / 4 + 1e-12) and np.all(y >= np.array(table.lower_bounds))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dr_type", ["uniform", "gaussian", "truncnorm"])
def test_sample_ids_across_the_2_32_boundary(dr_type, dtype):
    """The fast sampler loops keep Philox counter words 1..3 constant inside a 2^32-aligned segment of sample ids and
    split a tile that crosses a multiple of 2^32.  One launch over ids [2^32 - 777, 2^32 + 4000) must equal the two
    launches that stop / start at the boundary (neither of which crosses it), and, for the uniform law, the oracle."""
    table = random_envs.XI_TABLES["RandomHumanoid-v0"]
    lo = np.array([b[0] for b in table.search_bounds]); hi = np.array([b[1] for b in table.search_bounds])
    a, b = (lo, hi) if dr_type == "uniform" else ((lo + hi) / 2, (hi - lo) / 10)
    start = (1 << 32) - 777

    def draw(n, id0):
        s = _sampler("RandomHumanoid-v0", dr_type, a, b, seed=31)       # fresh sampler: call index 0 every time
        return s.sample_tasks_tensor(n, dtype=dtype, sample_id0=id0).cpu().numpy()
    whole = draw(4777, start)
    assert np.array_equal(whole[:777], draw(777, start)) and np.array_equal(whole[777:], draw(4000, 1 << 32))
    assert not np.array_equal(whole[777:1777], draw(1000, 0))          # the high id word is part of the key
    if dr_type == "uniform":
        npdt = np.float32 if dtype == torch.float32 else np.float64
        for i in (0, 776, 777, 778, 4776):
            assert np.array_equal(whole[i], c_oracle.xi_uniform(31, start + i, 0, lo, hi, purpose=c_oracle.PURPOSE_TASKS,
                                                                dtype=npdt)), i
