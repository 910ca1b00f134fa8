"""Host-side mirror of the reference's ``RandomEnv`` base class, sampling on the GPU.

Interface parity target: /root/reference/random_envs/random_env.py (``RandomEnv``, :4-259) -- same
method names, argument meaning, return types and exception types/messages.  What differs is where
the work happens: ``sample_task(s)`` launches the ``renv_dr_sample_*`` kernel (Philox + inverse-CDF
truncated normal) instead of looping over numpy/scipy scalar draws.

Deliberate deviations from the reference, all documented in DESIGN.md:
  * ``truncnorm`` works (the reference raises NameError: it never imports scipy, random_env.py:161).
  * draws come from a counter-based Philox stream keyed by (seed, sample index, call index), not from
    the process-global numpy state (random_env.py:151,180): same law, different stream.
  * the gaussian "three draws below 0.1" failure is detected on the device and raised on the host
    with the reference's message (random_env.py:186).
"""
import csv

import numpy as np

from . import _device, _lib
from .gym_compat import Env

SAMPLING_UNSET_MSG = ("sampling value of random env needs to be set before using sample_task() or "
                      "set_random_task(). Set it by uploading a DR distr.")
GAUSSIAN_FAIL_MSG = "Not all samples were above > 0.1 after 2 attempts"
INVALID_ACTION_MSG = ("an action outside Discrete(2) was passed to step(): the reference asserts '%r (%s) invalid' "
                      "(random_cartpole.py:173-174); detected on the device, the affected envs were pushed left")


def covariance_factor(cov):
    """A matrix F with F F^T = cov.  numpy's multivariate_normal (random_env.py:194) factors cov by SVD; any factor
    gives the same law.  Cholesky when cov is positive definite, symmetric eigen-decomposition otherwise (PSD)."""
    cov = np.asarray(cov, dtype=np.float64)
    if cov.ndim != 2 or cov.shape[0] != cov.shape[1]:
        raise ValueError("cov must be a square matrix")
    try:
        return np.linalg.cholesky(cov)
    except np.linalg.LinAlgError:
        w, v = np.linalg.eigh((cov + cov.T) / 2)
        return v * np.sqrt(np.clip(w, 0.0, None))


class RandomEnv(Env):
    """Superclass for all environments supporting Domain Randomization of dynamics parameters."""

    def __init__(self):
        self.sampling = None
        self.dr_training = False
        self.preferred_lr = None
        self.reward_threshold = None
        self.dyn_ind_to_name = None
        self._dr_seed = 0
        self._dr_calls = 0
        self._dr_violations = None     # device int64[NUM_COUNTERS] (include/renv.h `counters`), allocated on first GPU use

    # ---- hooks every env overrides (random_env.py:20-34) ----------------------------------------
    def get_search_bounds_mean(self, index):
        raise NotImplementedError

    def get_task_lower_bound(self, index):
        raise NotImplementedError

    def get_task(self):
        raise NotImplementedError

    def set_task(self, *task):
        raise NotImplementedError

    # ---- flags and small accessors (random_env.py:37-70) ----------------------------------------
    def set_random_task(self):
        self.set_task(*self.sample_task())

    def set_dr_training(self, flag):
        self.dr_training = flag

    def get_dr_training(self):
        return self.dr_training

    def set_endless(self, flag):
        self.endless = flag

    def get_endless(self):
        return self.endless

    def get_reward_threshold(self):
        return self.reward_threshold

    def dyn_index_to_name(self, index):
        assert self.dyn_ind_to_name is not None
        return self.dyn_ind_to_name[index]

    # ---- distribution state (random_env.py:72-143) ------------------------------------------------
    def set_dr_distribution(self, dr_type, distr):
        """dr_type in {uniform, truncnorm, gaussian, fullgaussian}; distr = [a0, b0, a1, b1, ...]."""
        if dr_type not in ("uniform", "truncnorm", "gaussian", "fullgaussian"):
            raise Exception("Unknown dr_type:" + str(dr_type))
        self.sampling = dr_type          # set first, like the reference's _set_* helpers (:102-127)
        self._on_distribution_change()
        if dr_type == "uniform":
            self._fill_interleaved(distr, self.min_task, self.max_task)
        elif dr_type in ("truncnorm", "gaussian"):
            self._fill_interleaved(distr, self.mean_task, self.stdev_task)
        else:
            self.mean_task[:] = distr["mean"]
            self.cov_task = np.copy(distr["cov"])

    # the reference's private per-type setters (random_env.py:102-127), kept for code that calls them directly
    def _set_udr_distribution(self, bounds):
        self.set_dr_distribution("uniform", bounds)

    def _set_truncnorm_distribution(self, bounds):
        self.set_dr_distribution("truncnorm", bounds)

    def _set_gaussian_distribution(self, bounds):
        self.set_dr_distribution("gaussian", bounds)

    def _set_fullgaussian_distribution(self, mean, cov):
        self.set_dr_distribution("fullgaussian", {"mean": mean, "cov": cov})

    @staticmethod
    def _fill_interleaved(distr, first, second):
        # range(len//2): a short list sets a prefix, a long one raises IndexError like the reference
        for i in range(len(distr) // 2):
            first[i] = distr[2 * i]
            second[i] = distr[2 * i + 1]

    def _on_distribution_change(self):
        """Subclasses holding device-side copies of the distribution refresh them here."""

    def get_dr_distribution(self):
        if self.sampling == "uniform":
            return self.min_task, self.max_task
        if self.sampling == "truncnorm":
            return self.mean_task, self.stdev_task
        if self.sampling == "gaussian":
            raise ValueError("Not implemented")
        return None

    def _task_len(self):
        # the reference uses len(self.get_task()); a vector env's get_task() is (N, task_dim), so prefer task_dim
        dim = getattr(self, "task_dim", None)
        return int(dim) if dim else len(self.get_task())

    def set_task_search_bounds(self):
        for i in range(self._task_len()):
            self.min_task[i], self.max_task[i] = self.get_search_bounds_mean(i)
        self._on_distribution_change()

    def get_task_search_bounds(self):
        dim = self._task_len()
        bounds = np.array([self.get_search_bounds_mean(i) for i in range(dim)], dtype=float).reshape(dim, 2)
        return bounds[:, 0].copy(), bounds[:, 1].copy()

    def denormalize_parameters(self, parameters):
        """Map parameters from the normalised [0, 4] space back to the search bounds (:205-220)."""
        assert parameters.shape[0] == self.task_dim
        lo, hi = self.get_task_search_bounds()
        return np.array((parameters * (hi - lo)) / 4 + lo)

    def load_dr_distribution_from_file(self, filename):
        """Two-line CSV: dr_type, then 2*task_dim floats (random_env.py:222-259)."""
        with open(filename, "r", encoding="utf-8") as fh:
            rows = csv.reader(fh, delimiter=",")
            dr_type = str(next(rows)[0])
            bounds = [float(col) for col in next(rows)]
        if len(bounds) != self.task_dim * 2:
            raise Exception("The file did not contain the right number of column values")
        if dr_type not in ("uniform", "truncnorm", "gaussian"):
            raise Exception("Filename is wrongly formatted: " + str(filename))
        self.set_dr_distribution(dr_type, bounds)

    # ---- sampling on the GPU (random_env.py:145-203) -----------------------------------------------
    def dr_config(self):
        """``renv_dr_cfg`` image of the current distribution (``None`` sampling -> DR_NONE)."""
        if self.sampling is None:
            return _lib.make_dr_cfg(None, None, None)
        if self.sampling == "uniform":
            a, b = self.min_task, self.max_task
        elif self.sampling in ("truncnorm", "gaussian"):
            a, b = self.mean_task, self.stdev_task
        elif self.sampling == "fullgaussian":
            lo, hi = self.get_task_search_bounds()
            return _lib.make_dr_cfg("fullgaussian", self.mean_task, lo, hi, covariance_factor(self.cov_task))
        else:
            raise Exception("Unknown dr_type:" + str(self.sampling))
        lb = [self.get_task_lower_bound(i) for i in range(len(a))] if self.sampling == "truncnorm" else None
        return _lib.make_dr_cfg(self.sampling, a, b, lb)

    def _violation_counter(self, device):
        t = _device.torch()
        if self._dr_violations is None or self._dr_violations.device != device:
            self._dr_violations = t.zeros(_lib.NUM_COUNTERS, dtype=t.int64, device=device)
        return self._dr_violations

    def check_dr_violations(self):
        """Raise the reference's gaussian failure if any device-side draw exhausted its 3 attempts."""
        if self._dr_violations is None:
            return
        gaussian, bad_action, order_timeout = (int(v) for v in self._dr_violations.tolist())
        if gaussian or bad_action or order_timeout:
            self._dr_violations.zero_()
        if order_timeout:
            raise RuntimeError("%d step CTAs timed out waiting for their tile's previous step (renv_cartpole_env.progress "
                               "was modified outside the step kernels, or steps sharing it ran unordered)" % order_timeout)
        if bad_action:
            raise AssertionError(INVALID_ACTION_MSG)
        if gaussian:
            raise Exception(GAUSSIAN_FAIL_MSG)

    def sample_tasks_tensor(self, num_tasks=1, dtype=None, device=None, out=None, sample_id0=0):
        """``sample_tasks`` without leaving the GPU: returns a (num_tasks, task_dim) CUDA tensor.  Row i is keyed by
        (seed_dr seed, sample_id0 + i, call index): a rank that owns global envs [id0, id0 + n) passes sample_id0=id0
        and draws exactly the rows the unsharded call would."""
        if self.sampling is None:
            raise ValueError(SAMPLING_UNSET_MSG)
        t = _device.torch()
        dev = _device.require_cuda(device if out is None else out.device)
        dtype = dtype or (out.dtype if out is not None else t.float32)
        cfg = self.dr_config()
        if out is None:
            out = t.empty((num_tasks, cfg.dim), dtype=dtype, device=dev)
        assert out.is_contiguous() and tuple(out.shape) == (num_tasks, cfg.dim) and out.dtype == dtype
        fn = {t.float32: "renv_dr_sample_f32", t.float64: "renv_dr_sample_f64"}[dtype]
        viol = self._violation_counter(dev)
        with t.cuda.device(dev):
            _lib.call(fn, _device.ptr(out), num_tasks, cfg, self._dr_seed, int(sample_id0), self._dr_calls, _device.ptr(viol),
                      _device.stream_ptr(dev))
        self._dr_calls = (self._dr_calls + 1) & 0xFFFFFFFF
        return out

    def sample_tasks(self, num_tasks=1):
        t = _device.torch()
        before = None
        if self._dr_violations is not None:         # violations of EARLIER launches are not this call's: keep them apart
            before = self._dr_violations.clone()
            self._dr_violations.zero_()
        try:
            out = self.sample_tasks_tensor(num_tasks, dtype=t.float64).cpu().numpy()
            self.check_dr_violations()
        finally:
            if before is not None:
                self._dr_violations += before
        return out

    def sample_task(self):
        """Sample random dynamics parameters -> float64 ndarray (task_dim,)."""
        return self.sample_tasks(1)[0]

    def seed_dr(self, seed):
        """Key of the Philox stream behind sample_task(s) (the reference uses the global numpy state)."""
        self._dr_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._dr_calls = 0


class TaskSampler(RandomEnv):
    """DR sampler for any env id of the suite, built from its xi table (no dynamics).

    Covers the "4-dim cartpole/hopper to 30-dim humanoid" sampler sweep: ``TaskSampler('RandomHumanoid-v0')``
    has the reference's ``task_dim``, ``dyn_ind_to_name``, lower bounds and search bounds, and its
    ``sample_task(s)`` follow ``RandomEnv.sample_task``.
    """

    def __init__(self, env_id):
        from .xi_tables import get_table
        super().__init__()
        table = get_table(env_id)
        self.env_id = env_id
        self._table = table
        self.task_dim = len(table.names)
        self.dyn_ind_to_name = dict(enumerate(table.names))
        self.min_task = np.zeros(self.task_dim)
        self.max_task = np.zeros(self.task_dim)
        self.mean_task = np.zeros(self.task_dim)
        self.stdev_task = np.zeros(self.task_dim)
        self.reward_threshold = table.reward_threshold
        self.preferred_lr = table.preferred_lr
        self._task = np.zeros(self.task_dim)

    def get_search_bounds_mean(self, index):
        return self._table.search_bounds[index]

    def get_task_lower_bound(self, index):
        return self._table.lower_bounds[index]

    def get_task(self):
        return self._task.copy()

    def set_task(self, *task):
        self._task = np.array(task, dtype=np.float64)
