"""``RandomCartPole-v0`` as a drop-in gym env: one env, float64, every step a CUDA kernel launch.

Interface parity target: /root/reference/random_envs/random_cartpole.py (``RandomCartPoleEnv``):
``reset() -> ndarray(4,) f64``, ``step(int) -> (ndarray(4,) f64, float, bool, dict)``,
``seed(seed) -> [seed]``, ``get_task()``, ``set_task(*xi)``, the bounds tables, the same attributes
(``gravity``, ``cart_mass``, ``pole_mass``, ``pole_length``, ``total_mass``, ``polemass_length`` ...)
and the same registration id / ``max_episode_steps=500`` (:291-296).

The physics is NOT computed here: the env owns an N=1 ``RandomCartPoleVecEnv`` (float64 parity path,
no auto-reset, TimeLimit left to the gym wrapper exactly as in the reference stack) and each ``step``
is one ``renv_cartpole_step_f64`` launch followed by a 56-byte device->host read.  That costs a
launch + sync per step, which is the price of the scalar gym API; throughput lives in the vector env.

Deliberate deviation (BASELINE.json north_star, README.md:9): with ``set_dr_training(True)`` ``reset``
resamples xi -- the reference CartPole forgets to (random_cartpole.py:226-229) although every MuJoCo
env of the suite does.  Pass ``resample_on_reset=False`` to get the reference's literal behaviour.
"""
import numpy as np

from . import _device, gym_compat
from .gym_compat import logger, spaces
from .random_env import RandomEnv
from .vector_env import (MAX_EPISODE_STEPS, NOMINAL_TASK, THETA_THRESHOLD_RADIANS, X_THRESHOLD, RandomCartPoleVecEnv,
                         _TABLE)


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
        self._core = RandomCartPoleVecEnv(1, dtype="float64", device=device, max_episode_steps=0, auto_reset=False,
                                          track_truncated=False, noisy=bool(noisy), noise_level=self.noise_level)
        self._pushed = None     # (state, task) last written to the device, to skip redundant uploads
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
        self._core.seed(seed)
        self.seed_dr(seed)
        return [seed]

    # ---- device round trip ---------------------------------------------------------------------------
    def _push(self):
        """Upload host-visible attributes the user may have assigned (env.state = ..., set_task)."""
        core = self._core
        core.kinematics_integrator = self.kinematics_integrator
        core.noise_level = float(self.noise_level)
        task = (float(self.gravity), float(self.cart_mass), float(self.pole_mass), float(self.pole_length))
        state = tuple(float(v) for v in self.state)
        beyond = -1 if self.steps_beyond_done is None else int(self.steps_beyond_done)
        if self._pushed != (state, task, beyond):
            core.set_task(*task)
            core.set_state(np.array(state).reshape(1, 4))
            core.steps_beyond_done.fill_(beyond)

    def step(self, action):
        err_msg = "%r (%s) invalid" % (action, type(action))
        assert self.action_space.contains(action), err_msg
        if self.state is None:
            raise TypeError("cannot unpack non-iterable NoneType object")   # the reference's failure before reset()
        self._push()
        was_beyond = self.steps_beyond_done
        t = _device.torch()
        obs, reward, done, _ = self._core.step(t.tensor([int(action)], dtype=t.uint8))
        packed = t.cat([self._core.state.reshape(-1), reward.reshape(-1), done.reshape(-1).to(t.float64),
                        self._core.steps_beyond_done.to(t.float64), obs.reshape(-1)]).cpu().numpy()
        self.state = tuple(float(v) for v in packed[:4])
        reward, done, beyond = float(packed[4]), bool(packed[5]), int(packed[6])
        observation = packed[7:11].copy() if self.noisy else np.array(self.state)
        self.steps_beyond_done = None if beyond < 0 else beyond
        if was_beyond == 0 and done:
            logger.warn("You are calling 'step()' even though this environment has already returned done = True. "
                        "You should always call 'reset()' once you receive 'done = True' -- any further steps are "
                        "undefined behavior.")
        self._pushed = (self.state, (float(self.gravity), float(self.cart_mass), float(self.pole_mass),
                                     float(self.pole_length)), beyond)
        return observation, reward, done, {}

    def reset(self):
        if self.dr_training and self.resample_on_reset and self.sampling is not None:
            self.set_random_task()
        # s0 ~ U(-0.05, 0.05)^4 drawn by the reset kernel (Philox keyed by seed / episode)
        core = self._core
        core.kinematics_integrator = self.kinematics_integrator
        core.noise_level = float(self.noise_level)
        obs = core.reset()
        self.state = tuple(float(v) for v in core.state.reshape(-1).cpu().numpy())
        observation = obs.reshape(-1).cpu().numpy().copy() if self.noisy else np.array(self.state)
        self.steps_beyond_done = None
        core.set_task(float(self.gravity), float(self.cart_mass), float(self.pole_mass), float(self.pole_length))
        self._pushed = (self.state, (float(self.gravity), float(self.cart_mass), float(self.pole_mass),
                                     float(self.pole_length)), -1)
        return observation

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
