"""``RandomCartPole-v0`` as a drop-in gym env: one env, float64, every step a CUDA kernel launch.

Interface parity target: /root/reference/random_envs/random_cartpole.py (``RandomCartPoleEnv``):
``reset() -> ndarray(4,) f64``, ``step(int) -> (ndarray(4,) f64, float, bool, dict)``,
``seed(seed) -> [seed]``, ``get_task()``, ``set_task(*xi)``, the bounds tables, the same attributes
(``gravity``, ``cart_mass``, ``pole_mass``, ``pole_length``, ``total_mass``, ``polemass_length`` ...)
and the same registration id / ``max_episode_steps=500`` (:291-296).

The physics is NOT computed here.  By default the env is served by a RESIDENT one-warp kernel
(``renv_cartpole_scalar_serve``, csrc/renv_scalar_server.cuh): ``step`` writes one request word into pinned,
device-mapped host memory and spins on the acknowledgement; state, xi and ``steps_beyond_done`` live in the kernel's
registers between calls (float64 parity arithmetic, no auto-reset, TimeLimit left to the gym wrapper exactly as in the
reference stack).  No launch, no stream synchronise: ~4 us per step instead of ~28 us.  The kernel is a lease: after
``RENV_SCALAR_LEASE_US`` (default 300) microseconds without a request it parks the env in device memory and exits,
and the next call relaunches it.  ``RENV_SCALAR_RESIDENT=0`` selects the older path, one
``renv_cartpole_step_f64`` launch + stream synchronise per call on mapped buffers.

Deliberate deviation (BASELINE.json north_star, README.md:9): with ``set_dr_training(True)`` ``reset``
resamples xi -- the reference CartPole forgets to (random_cartpole.py:226-229) although every MuJoCo
env of the suite does.  Pass ``resample_on_reset=False`` to get the reference's literal behaviour.
"""
import ctypes
import struct
import math

import numpy as np

from . import _device, _lib, gym_compat
from .gym_compat import logger, spaces
from .random_env import GAUSSIAN_FAIL_MSG, RandomEnv
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


class _ResidentScalarCore:
    """One float64 env inside the resident server kernel (include/renv.h renv_cartpole_scalar_serve)."""

    # byte offsets of struct renv_scalar_ctrl
    OFF_REQUEST, OFF_ARG, OFF_ARG_U64, OFF_STATE, OFF_OBS, OFF_XI, OFF_REWARD, OFF_DONE, OFF_BEYOND, OFF_VIOL, OFF_ACK, \
        OFF_EXITED, OFF_NEXT, NEXT_STRIDE, OFF_NEXT_SEQ, SIZE = 0, 64, 192, 256, 288, 320, 352, 360, 364, 368, 376, 380, \
        448, 128, 704, 768
    OP_RESET, OP_SET_STATE, OP_SET_XI, OP_CONFIG, OP_EXIT, OP_AHEAD_STEP = 2, 3, 4, 5, 6, 8

    def __init__(self, device, noisy):
        import os
        t = _device.torch()
        self.device = _device.require_cuda(device)
        self.lib = _lib.load()
        self.noisy = bool(noisy)
        self._pin = t.zeros(1024, dtype=t.uint8).pin_memory()              # renv_scalar_ctrl, device-mapped host memory
        raw = self._pin.numpy()
        u32 = lambda off, n=1: raw[off:off + 4 * n].view(np.uint32)          # noqa: E731
        f64 = lambda off, n: raw[off:off + 8 * n].view(np.float64)           # noqa: E731
        self.request, self.ack, self.exited = u32(self.OFF_REQUEST), u32(self.OFF_ACK), u32(self.OFF_EXITED)
        self.arg, self.arg_u64 = f64(self.OFF_ARG, 16), raw[self.OFF_ARG_U64:self.OFF_ARG_U64 + 64].view(np.uint64)
        self.out_state, self.out_obs, self.out_xi = f64(self.OFF_STATE, 4), f64(self.OFF_OBS, 4), f64(self.OFF_XI, 4)
        self.out_reward = f64(self.OFF_REWARD, 1)
        self.out_done = raw[self.OFF_DONE:self.OFF_DONE + 8].view(np.int32)         # done, beyond
        self.out_viol = u32(self.OFF_VIOL)
        # look-ahead block: what step(0) / step(1) would return from the kernel's current state (renv_scalar_ctrl.next)
        self.next_seq = u32(self.OFF_NEXT_SEQ)
        nx = lambda a: self.OFF_NEXT + a * self.NEXT_STRIDE                   # noqa: E731
        self.next_state = [f64(nx(a), 4) for a in (0, 1)]
        self.next_obs = [f64(nx(a) + 32, 4) for a in (0, 1)]
        self.next_reward = [f64(nx(a) + 64, 1) for a in (0, 1)]
        self.next_done = [raw[nx(a) + 72:nx(a) + 80].view(np.int32) for a in (0, 1)]      # done, beyond
        self.lookahead = os.environ.get("RENV_SCALAR_LOOKAHEAD", "1") != "0"
        # the words the step path polls / writes, as plain Python ints (a memoryview item costs a third of a numpy one)
        self._words = memoryview(raw).cast("B").cast("I")
        self._raw = raw
        self._unpack_outcome = struct.Struct("<4d4ddii").unpack_from     # state, obs, reward, done, beyond
        self.obs_now = None                    # Noisy variant: the observation of the last step / reset
        with t.cuda.device(self.device):
            self._save = t.zeros(_lib.SCALAR_SAVE_BYTES, dtype=t.uint8, device=self.device)
            self._stream = t.cuda.Stream(device=self.device)
        self._ctrl_ptr, self._save_ptr = ctypes.c_void_p(self._pin.data_ptr()), ctypes.c_void_p(self._save.data_ptr())
        self._lease_ns = int(float(os.environ.get("RENV_SCALAR_LEASE_US", "300")) * 1000)
        self._lease_id, self._running, self._seq, self._pending = 0, False, 0, False
        self.tick = 0
        self.state_tuple, self.xi_tuple, self.beyond = None, None, None
        self.config_key = None

    # ---- the doorbell -------------------------------------------------------------------------------------
    def _launch(self):
        t = _device.torch()
        self._lease_id = (self._lease_id + 1) & 0xFFFFFFFF or 1
        with t.cuda.device(self.device):
            rc = self.lib.renv_cartpole_scalar_serve(self._ctrl_ptr, self._save_ptr, self._lease_id, self._lease_ns,
                                                     ctypes.c_void_p(self._stream.cuda_stream))
        if rc != _lib.OK:
            raise _lib.RenvError("renv_cartpole_scalar_serve", rc, _lib.strerror(rc))
        self._running = True

    W_REQUEST, W_ACK, W_EXITED, W_NEXT_SEQ = OFF_REQUEST // 4, OFF_ACK // 4, OFF_EXITED // 4, OFF_NEXT_SEQ // 4

    def _wait(self, word, seq):
        """Spin until the kernel stored `seq` into 32-bit word `word` (W_ACK or W_NEXT_SEQ); bounded: relaunches an
        expired lease (the new instance finds the request still pending), raises after ~5 s."""
        w, spins = self._words, 0
        while w[word] != seq:
            spins += 1
            if not spins & 0x3FF:                     # every 1024 polls (~50 us): did the lease expire under us?
                if w[self.W_EXITED] == self._lease_id and w[word] != seq:
                    self._launch()
                elif spins > 50_000_000:
                    raise RuntimeError("the resident scalar-env kernel did not answer request %d" % seq)

    def _ring(self, op):
        """Write one request word.  At most one request is outstanding: the caller has waited for the previous one."""
        w = self._words
        if not self._running or w[self.W_EXITED] == self._lease_id:
            self._launch()
        seq = self._seq = (self._seq + 1) & 0xFFFFFF
        w[self.W_REQUEST] = (seq << 8) | op
        return seq

    def _call(self, op):
        """Ring one request in and wait for its acknowledgement."""
        if self._pending:                             # a look-ahead step is still in flight: let the kernel take it first
            self._wait(self.W_NEXT_SEQ, self._seq)    # (look-ahead steps are not acknowledged, their successors are published)
            self._pending = False
        self._wait(self.W_ACK, self._ring(op))

    def close(self):
        if self._running and (self._pending or self.exited[0] != self._lease_id):
            try:
                self._call(self.OP_EXIT)
            except Exception:  # noqa: BLE001
                pass
        self._running = False

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- requests -------------------------------------------------------------------------------------------
    def _fetch(self, with_xi):
        self.state_tuple = tuple(self.out_state.tolist())
        b = int(self.out_done[1])
        self.beyond = None if b < 0 else b
        if with_xi:
            self.xi_tuple = tuple(self.out_xi.tolist())

    def configure(self, seed, integrator, noise_level, dr_cfg):
        """dr_cfg: _lib.DrCfg image of the distribution (or None): converted like renv_abi.cu to_cfg4 does."""
        a = self.arg
        a[:] = 0.0
        a[0] = math.sqrt(noise_level)
        dr_type = 0
        if dr_cfg is not None and dr_cfg.dr_type in (_lib.DR_UNIFORM, _lib.DR_TRUNCNORM, _lib.DR_GAUSSIAN):
            dr_type = dr_cfg.dr_type
            for k in range(4):
                a[1 + k] = dr_cfg.a[k]
                a[5 + k] = dr_cfg.b[k] - dr_cfg.a[k] if dr_type == _lib.DR_UNIFORM else dr_cfg.b[k]
                a[9 + k] = dr_cfg.lb[k] if dr_type == _lib.DR_TRUNCNORM else (0.1 if dr_type == _lib.DR_GAUSSIAN else 0.0)
        u = self.arg_u64
        u[0], u[1], u[2], u[3] = int(seed) & 0xFFFFFFFFFFFFFFFF, 1 if integrator == _lib.EULER else 0, dr_type, int(self.noisy)
        u[4] = self.tick                  # the kernel counts the step clock itself from here (look-ahead observations)
        u[5] = int(self.lookahead)
        self._call(self.OP_CONFIG)
        self._fetch(True)

    def set_state(self, state, beyond):
        self.arg[0:4] = state
        self.arg_u64[0] = np.uint64(np.int64(-1 if beyond is None else int(beyond)).astype(np.uint64))
        self._call(self.OP_SET_STATE)
        self._fetch(False)

    def set_xi(self, xi):
        self.arg[0:4] = xi
        self._call(self.OP_SET_XI)
        self._fetch(True)

    def reset(self, resample, dr_call, dr_seed):
        u = self.arg_u64
        u[0], u[1], u[2], u[3] = self.tick, int(resample), int(dr_call), int(dr_seed)
        self._call(self.OP_RESET)
        self.tick += 1
        self._fetch(True)
        if self.noisy:
            self.obs_now = self.out_obs.copy()
        return int(self.out_viol[0])

    def step(self, action):
        if self.lookahead:
            # The kernel has published both outcomes of the next step from its current state (next_seq == the last
            # request's seq once it has).  Take ours and ring the step in WITHOUT waiting for it: the PCIe round trip
            # overlaps whatever the caller does before its next call.
            seq = self._seq
            if self._words[self.W_NEXT_SEQ] != seq:
                self._wait(self.W_NEXT_SEQ, seq)
            # next_seq == seq: the kernel has consumed request seq and next[action] is what this step returns
            out = self._unpack_outcome(self._raw, self.OFF_NEXT + action * self.NEXT_STRIDE)
            self._ring(self.OP_AHEAD_STEP + action)
            self._pending = True
            self.tick += 1
            self.state_tuple = out[0:4]
            self.beyond = None if out[10] < 0 else out[10]
            if self.noisy:
                self.obs_now = np.array(out[4:8])
            return out[8], bool(out[9])
        self._call(action)
        self.tick += 1
        self.state_tuple = tuple(self.out_state.tolist())
        b = int(self.out_done[1])
        self.beyond = None if b < 0 else b
        if self.noisy:
            self.obs_now = self.out_obs.copy()
        return float(self.out_reward[0]), bool(self.out_done[0])


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
        import os
        self._resident = os.environ.get("RENV_SCALAR_RESIDENT", "1") != "0"
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
            if self._resident:
                self._core_obj.config_key = None      # re-sent with the next request
                self._core_obj.tick = 0               # the reference rebuilds np_random: seed(s); reset() repeats
            else:
                self._core_obj.seed(seed)
                self._core_obj.tick = 0
        self.seed_dr(seed)
        return [seed]

    # ---- device round trip ---------------------------------------------------------------------------
    @property
    def _core(self):
        """The device-side env, created on first use (constructing the env needs no GPU; computing does)."""
        if self._core_obj is None:
            if self._resident:
                self._core_obj = _ResidentScalarCore(*self._core_args)
            else:
                self._core_obj = _MappedScalarCore(*self._core_args)
                self._core_obj.seed(self._core_seed)
        return self._core_obj

    def _integrator(self):
        return _lib.EULER if self.kinematics_integrator == "euler" else _lib.SEMI_IMPLICIT

    def _sync_resident(self, core, want_dr):
        """Bring the kernel's copy in line with whatever the user assigned on the Python object since the last call
        (env.state = ..., set_task, steps_beyond_done, kinematics_integrator, noise_level, the DR distribution)."""
        dr_cfg = None
        if want_dr and self.sampling in ("uniform", "truncnorm", "gaussian"):
            dr_cfg = self.dr_config()
        key = (self._core_seed, self.kinematics_integrator, float(self.noise_level),
               None if dr_cfg is None else bytes(dr_cfg)[:8 + 3 * 8 * 4 * 8])
        if key != core.config_key:
            core.configure(self._core_seed, self._integrator(), float(self.noise_level), dr_cfg)
            core.config_key = key
        xi = (self.gravity, self.cart_mass, self.pole_mass, self.pole_length)
        if xi != core.xi_tuple:
            core.set_xi(xi)

    def step(self, action):
        assert self.action_space.contains(action), "%r (%s) invalid" % (action, type(action))   # :173-174
        if self.state is None:
            raise TypeError("cannot unpack non-iterable NoneType object")   # the reference's failure before reset()
        core = self._core
        was_beyond = self.steps_beyond_done
        if self._resident:
            self._sync_resident(core, False)
            if self.state is not core.state_tuple or was_beyond != core.beyond:
                core.set_state(self.state, was_beyond)
            reward, done = core.step(int(action))
            self.state = core.state_tuple
            self.steps_beyond_done = core.beyond
            observation = core.obs_now if self.noisy else np.array(self.state)
        else:
            # host-visible attributes the user may have assigned (env.state = ..., set_task, steps_beyond_done): plain
            # stores into the mapped buffers, no upload
            core.state[:, 0] = self.state
            core.xi[0] = (self.gravity, self.cart_mass, self.pole_mass, self.pole_length)
            core.beyond[0] = -1 if was_beyond is None else int(was_beyond)
            core.step(int(action), self._integrator(), float(self.noise_level))
            self.state = tuple(float(v) for v in core.state[:, 0])
            reward, done, beyond = float(core.reward[0]), bool(core.done[0]), int(core.beyond[0])
            self.steps_beyond_done = None if beyond < 0 else beyond
            observation = core.obs[:, 0].copy() if self.noisy else np.array(self.state)
        if was_beyond == 0 and done:
            logger.warn("You are calling 'step()' even though this environment has already returned done = True. "
                        "You should always call 'reset()' once you receive 'done = True' -- any further steps are "
                        "undefined behavior.")
        return observation, reward, done, {}

    def reset(self):
        resample = bool(self.dr_training and self.resample_on_reset and self.sampling is not None)
        core = self._core
        if self._resident:
            in_kernel = resample and self.sampling in ("uniform", "truncnorm", "gaussian")
            if resample and not in_kernel:
                self.set_random_task()                  # fullgaussian: sampled by its own kernel, then uploaded below
            self._sync_resident(core, in_kernel)
            # s0 ~ U(-0.05, 0.05)^4 (Philox keyed by seed / tick) and, for the per-dim laws, xi = sample_task(): the
            # very draws RandomEnv.sample_task() would return for this call index, made inside the resident kernel
            viol = core.reset(in_kernel, self._dr_calls, self._dr_seed)
            if in_kernel:
                self._dr_calls = (self._dr_calls + 1) & 0xFFFFFFFF
                self.set_task(*core.xi_tuple)
                if viol:
                    raise Exception(GAUSSIAN_FAIL_MSG)
            self.state = core.state_tuple
            self.steps_beyond_done = None
            return core.obs_now if self.noisy else np.array(self.state)
        if resample:
            self.set_random_task()
        # s0 ~ U(-0.05, 0.05)^4 drawn by the reset kernel (Philox keyed by seed / tick)
        core.reset(float(self.noise_level))
        self.state = tuple(float(v) for v in core.state[:, 0])
        self.steps_beyond_done = None
        return core.obs[:, 0].copy() if self.noisy else np.array(self.state)

    def render(self, mode="human"):
        raise NotImplementedError("rendering (pyglet viewer, random_cartpole.py:231-288) is out of scope")

    def close(self):
        self.viewer = None
        if self._resident and self._core_obj is not None:
            self._core_obj.close()


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
