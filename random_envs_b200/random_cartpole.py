"""``RandomCartPole-v0`` as a drop-in gym env: one env, float64, every step a CUDA kernel launch.

Interface parity target: /root/reference/random_envs/random_cartpole.py (``RandomCartPoleEnv``):
``reset() -> ndarray(4,) f64``, ``step(int) -> (ndarray(4,) f64, float, bool, dict)``,
``seed(seed) -> [seed]``, ``get_task()``, ``set_task(*xi)``, the bounds tables, the same attributes
(``gravity``, ``cart_mass``, ``pole_mass``, ``pole_length``, ``total_mass``, ``polemass_length`` ...)
and the same registration id / ``max_episode_steps=500`` (:291-296).

The physics is NOT computed here: each ``step`` is one ``renv_cartpole_step_f64`` launch (float64 parity path, N = 1,
no auto-reset, TimeLimit left to the gym wrapper exactly as in the reference stack).  The env's few hundred bytes
(state, xi, action, reward, done, steps_beyond_done) live in PINNED host memory, which unified addressing maps into
the GPU's address space at the same pointer: the kernel reads the action and writes its results straight across
PCIe, so a step is one launch + one stream synchronize with no memcpy call at all (~15 us instead of ~105 us through
device buffers and tensor round trips).  That is still the price of the scalar gym API; throughput lives in the
vector env.

Deliberate deviation (BASELINE.json north_star, README.md:9): with ``set_dr_training(True)`` ``reset``
resamples xi -- the reference CartPole forgets to (random_cartpole.py:226-229) although every MuJoCo
env of the suite does.  Pass ``resample_on_reset=False`` to get the reference's literal behaviour.
"""
import ctypes
import math

import numpy as np

from . import _device, _lib, gym_compat
from .gym_compat import logger, spaces
from .random_env import RandomEnv
from .vector_env import MAX_EPISODE_STEPS, NOMINAL_TASK, THETA_THRESHOLD_RADIANS, X_THRESHOLD, _TABLE


class _MappedScalarCore:
    """One float64 env whose buffers are pinned host tensors used BY THE KERNELS through their (identical) device
    address: cudaHostAlloc'd memory is mapped under unified addressing, so no copy is ever enqueued."""

    LD = 2          # f64 rows need ld % 2 == 0

    def __init__(self, device, noisy):
        t = _device.torch()
        self.device = _device.require_cuda(device)
        _lib.load()
        pin = lambda shape, dtype: t.zeros(shape, dtype=dtype).pin_memory()      # noqa: E731
        self._keep = dict(state=pin((4, self.LD), t.float64), obs=pin((4, self.LD), t.float64), xi=pin((self.LD, 4), t.float64),
                          elapsed=pin(4, t.int32), beyond=pin(4, t.int32), reward=pin(2, t.float64), done=pin(16, t.uint8),
                          action=pin(16, t.uint8))
        k = self._keep
        self.state, self.obs, self.xi = k["state"].numpy(), k["obs"].numpy(), k["xi"].numpy()
        self.beyond, self.reward, self.done, self.action = (k["beyond"].numpy(), k["reward"].numpy(), k["done"].numpy(),
                                                            k["action"].numpy())
        self.beyond[:] = -1
        self.viol = t.zeros(_lib.NUM_COUNTERS, dtype=t.int64, device=self.device)
        env = _lib.CartpoleEnv()
        env.state, env.xi, env.elapsed = k["state"].data_ptr(), k["xi"].data_ptr(), k["elapsed"].data_ptr()
        env.episode, env.beyond = None, k["beyond"].data_ptr()
        env.n, env.ld, env.env_id0, env.seed = 1, self.LD, 0, 0
        self.env = env
        self.noise = _lib.ObsNoise()
        self.noise.obs, self.noise.std = k["obs"].data_ptr(), 0.0
        self.noisy = bool(noisy)
        self.tick = 0
        lib = _lib.load()
        self._reset = lib.renv_cartpole_reset_noisy_f64 if noisy else lib.renv_cartpole_reset_f64
        self._step = lib.renv_cartpole_step_noisy_f64 if noisy else lib.renv_cartpole_step_f64
        self._head = (ctypes.byref(env), ctypes.byref(self.noise)) if noisy else (ctypes.byref(env),)
        self._ptr = {n: ctypes.c_void_p(k[n].data_ptr()) for n in ("action", "reward", "done")}
        self._viol_ptr = ctypes.c_void_p(self.viol.data_ptr())

    def seed(self, seed):
        self.env.seed = int(seed) & 0xFFFFFFFFFFFFFFFF

    def _run(self, fn, name, *args):
        t = _device.torch()
        stream = t.cuda.current_stream(self.device)
        if t.cuda.current_device() == self.device.index:
            rc = fn(*self._head, *args, ctypes.c_void_p(stream.cuda_stream))
        else:
            with t.cuda.device(self.device):
                rc = fn(*self._head, *args, ctypes.c_void_p(stream.cuda_stream))
        if rc != _lib.OK:
            raise _lib.RenvError(name, rc, _lib.strerror(rc))
        stream.synchronize()            # the kernel's stores to the mapped buffers are visible after this
        self.tick += 1

    def reset(self, noise_level):
        """RandomCartPoleEnv.reset at the current tick (no DR here: the wrapper resamples on the host side)."""
        self.noise.std = math.sqrt(noise_level)
        self._run(self._reset, "renv_cartpole_reset_f64", None, self.tick, None, self._viol_ptr)

    def step(self, action, integrator, noise_level):
        self.action[0] = action
        self.noise.std = math.sqrt(noise_level)
        self._run(self._step, "renv_cartpole_step_f64", self._ptr["action"], self._ptr["reward"], self._ptr["done"], None,
                  integrator, 0, 0, self.tick, None, self._viol_ptr)


class RandomCartPoleEnv(RandomEnv):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 50}

    def __init__(self, noisy=False, device=None, resample_on_reset=True):
        RandomEnv.__init__(self)
        # `noisy` / `noise_level` as in the suite's MuJoCo envs (e.g. jinja/random_hopper.py:16,21-28,107-108); the
        # reference has no noisy cart-pole, RandomCartPoleNoisy-v0 is this framework's sibling of RandomHopperNoisy-v0
        self.noisy = noisy
        self.noise_level = 1e-4
        self.gravity, self.cart_mass, self.pole_mass, self.pole_length = NOMINAL_TASK
        self.total_mass = self.pole_mass + self.cart_mass
        self.polemass_length = self.pole_mass * self.pole_length   # never refreshed by set_task (:79, :157-166)
        self.force_mag = 10.0
        self.tau = 0.02
        self.kinematics_integrator = "euler"
        self.theta_threshold_radians = THETA_THRESHOLD_RADIANS
        self.x_threshold = X_THRESHOLD
        high = np.array([self.x_threshold * 2, np.finfo(np.float32).max, self.theta_threshold_radians * 2,
                         np.finfo(np.float32).max], dtype=np.float32)
        self.action_space = spaces.Discrete(2)
        self.observation_space = spaces.Box(-high, high, dtype=np.float32)
        self.viewer = None
        self.state = None
        self.steps_beyond_done = None
        self.dyn_ind_to_name = dict(enumerate(_TABLE.names))
        self.original_task = np.array(NOMINAL_TASK)
        self.task_dim = 4
        self.min_task = np.zeros(4)
        self.max_task = np.zeros(4)
        self.mean_task = np.zeros(4)
        self.stdev_task = np.zeros(4)
        self.reward_threshold = 500
        self.resample_on_reset = resample_on_reset
        self._core_obj, self._core_args, self._core_seed = None, (device, bool(noisy)), 0    # buffers are created on first use
        self.seed()

    # ---- tables ------------------------------------------------------------------------------------
    def get_search_bounds_mean(self, index):
        return _TABLE.search_bounds[index]

    def get_task_lower_bound(self, index):
        return _TABLE.lower_bounds[index]

    def get_task(self):
        return np.array([self.gravity, self.cart_mass, self.pole_mass, self.pole_length])

    def set_task(self, *task):
        self.gravity, self.cart_mass, self.pole_mass, self.pole_length = task[0], task[1], task[2], task[3]
        self.total_mass = self.pole_mass + self.cart_mass

    def seed(self, seed=None):
        self.np_random, seed = gym_compat.utils.seeding.np_random(seed)
        self._core_seed = seed
        if self._core_obj is not None:
            self._core_obj.seed(seed)
        self.seed_dr(seed)
        return [seed]

    # ---- device round trip ---------------------------------------------------------------------------
    @property
    def _core(self):
        """The mapped-memory env, created on first use (constructing the env needs no GPU; computing does)."""
        if self._core_obj is None:
            self._core_obj = _MappedScalarCore(*self._core_args)
            self._core_obj.seed(self._core_seed)
        return self._core_obj

    def step(self, action):
        err_msg = "%r (%s) invalid" % (action, type(action))
        assert self.action_space.contains(action), err_msg
        if self.state is None:
            raise TypeError("cannot unpack non-iterable NoneType object")   # the reference's failure before reset()
        core = self._core
        # host-visible attributes the user may have assigned (env.state = ..., set_task, steps_beyond_done): plain
        # stores into the mapped buffers, no upload
        core.state[:, 0] = self.state
        core.xi[0] = (self.gravity, self.cart_mass, self.pole_mass, self.pole_length)
        was_beyond = self.steps_beyond_done
        core.beyond[0] = -1 if was_beyond is None else int(was_beyond)
        core.step(int(action), _lib.EULER if self.kinematics_integrator == "euler" else _lib.SEMI_IMPLICIT,
                  float(self.noise_level))
        self.state = tuple(float(v) for v in core.state[:, 0])
        reward, done, beyond = float(core.reward[0]), bool(core.done[0]), int(core.beyond[0])
        self.steps_beyond_done = None if beyond < 0 else beyond
        if was_beyond == 0 and done:
            logger.warn("You are calling 'step()' even though this environment has already returned done = True. "
                        "You should always call 'reset()' once you receive 'done = True' -- any further steps are "
                        "undefined behavior.")
        observation = core.obs[:, 0].copy() if self.noisy else np.array(self.state)
        return observation, reward, done, {}

    def reset(self):
        if self.dr_training and self.resample_on_reset and self.sampling is not None:
            self.set_random_task()
        # s0 ~ U(-0.05, 0.05)^4 drawn by the reset kernel (Philox keyed by seed / tick)
        core = self._core
        core.reset(float(self.noise_level))
        self.state = tuple(float(v) for v in core.state[:, 0])
        self.steps_beyond_done = None
        return core.obs[:, 0].copy() if self.noisy else np.array(self.state)

    def render(self, mode="human"):
        raise NotImplementedError("rendering (pyglet viewer, random_cartpole.py:231-288) is out of scope")

    def close(self):
        self.viewer = None


gym_compat.register(
    id="RandomCartPole-v0",
    entry_point="%s:RandomCartPoleEnv" % __name__,
    max_episode_steps=MAX_EPISODE_STEPS,
    kwargs={},
)
# the suite's naming for the observation-noise siblings (jinja/random_hopper.py:161-166 RandomHopperNoisy-v0)
gym_compat.register(
    id="RandomCartPoleNoisy-v0",
    entry_point="%s:RandomCartPoleEnv" % __name__,
    max_episode_steps=MAX_EPISODE_STEPS,
    kwargs={"noisy": True},
)
