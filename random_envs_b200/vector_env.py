"""Vectorised, tensor-in/tensor-out RandomCartPole: N envs resident in HBM, one kernel per call.

Same method names as the reference's single env (random_cartpole.py / random_env.py): ``reset``,
``step``, ``seed``, ``get_task``, ``set_task``, ``set_random_task``, ``set_dr_distribution``,
``set_dr_training`` ...; arguments and results are batched torch CUDA tensors:

    reset()        -> obs (N, 4)
    step(actions)  -> obs (N, 4), reward (N,), done (N,) bool, info {'TimeLimit.truncated': (N,) bool}
    get_task()     -> (N, 4)            set_task(Tensor(N, 4) | g, m_c, m_p, l)
    rollout(w, b, K) -> None            (K fused steps of the linear policy a = [w.s + b > 0])
    episode_stats()  -> dict            (returns of the episodes finished inside rollouts)

``step`` fuses what the reference stack does in three Python layers: ``RandomCartPoleEnv.step``
(random_cartpole.py:172-224), gym 0.21 ``TimeLimit.step`` (max_episode_steps=500, :294) and gym 0.21
``SyncVectorEnv`` auto-reset, including the DR resample on reset when ``dr_training`` is on
(README.md:9; random_env.py:37-39).  The returned tensors are views of the env's own buffers (the
reference's ``obs`` likewise aliases ``state``): they are overwritten by the next call.

State is stored structure-of-arrays ``(4, ld)`` so that the kernels read and write 128-bit coalesced
vectors (``obs`` is the transposed ``(N, 4)`` view); xi is stored one row per env, ``(N, 4)``, which is
what ``get_task`` returns.
"""
import ctypes
import math

import numpy as np

from . import _device, _lib
from .random_env import RandomEnv
from .xi_tables import get_table

_TABLE = get_table("RandomCartPole-v0")
NOMINAL_TASK = (9.8, 1.0, 0.1, 0.5)                 # random_cartpole.py:74-78
THETA_THRESHOLD_RADIANS = 12 * 2 * math.pi / 360   # :85
X_THRESHOLD = 2.4                                  # :86
MAX_EPISODE_STEPS = 500                            # :294


def _round_up(x, m):
    return (x + m - 1) // m * m


class RandomCartPoleVecEnv(RandomEnv):
    """N domain-randomised cart-poles stepped by hand-written sm_100a kernels."""

    def __init__(self, num_envs, dtype="float32", device=None, seed=0, env_id0=0,
                 max_episode_steps=MAX_EPISODE_STEPS, auto_reset=True, kinematics_integrator="euler",
                 track_truncated=True, validate_actions=True, track_episodes=True, noisy=False, noise_level=1e-4,
                 lean=False, tile_ordering="auto", pack_host_flags=True):
        RandomEnv.__init__(self)
        if num_envs <= 0:
            raise ValueError("num_envs must be positive")
        self.num_envs = int(num_envs)
        # step_host_*: done / truncated cross PCIe as bits (renv_pack_flags_u8) and are unpacked on the host
        self.pack_host_flags = bool(pack_host_flags)
        self._dtype_name = str(dtype).replace("torch.", "")
        if self._dtype_name not in ("float32", "float64"):
            raise ValueError("dtype must be float32 or float64")
        self._device_arg = device
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.env_id0 = int(env_id0)
        self.max_episode_steps = 0 if max_episode_steps is None else int(max_episode_steps)
        self.auto_reset = bool(auto_reset)
        self.kinematics_integrator = kinematics_integrator
        self.track_truncated = bool(track_truncated)
        self.validate_actions = bool(validate_actions)
        self.track_episodes = bool(track_episodes)
        # the suite's "Noisy" variants (jinja/random_hopper.py:16,28,107-108): obs = state + sqrt(noise_level) * N(0, I)
        self.noisy = bool(noisy)
        self.noise_level = float(noise_level)
        if self.noise_level < 0:
            raise ValueError("noise_level must be >= 0")
        # lean step (include/renv.h renv_cartpole_step_lean_f32): uint16 TimeLimit counter, no reward store -- 54 instead
        # of 62 bytes of HBM traffic per env-step; state / done / truncated bit-identical
        self.lean = bool(lean)
        if self.lean and (self._dtype_name != "float32" or not self.auto_reset or self.noisy
                          or not 0 < self.max_episode_steps <= 65535):
            raise ValueError("lean=True needs float32, auto_reset, no observation noise and 0 < max_episode_steps <= 65535")

        # step-to-step ordering per 1024-env tile instead of per launch (include/renv.h renv_cartpole_env.progress):
        # consecutive step() launches of one stream overlap; False = plain stream order between launches.  "auto": on
        # up to 2^21 envs, where the launch-to-launch bubble is >= 10 % of a step; beyond that (and for steps issued
        # from several streams / parallel graph branches) the per-CTA ticket round trip costs more than it saves
        self.tile_ordering = (self.num_envs <= (1 << 21)) if tile_ordering == "auto" else bool(tile_ordering)

        self.dyn_ind_to_name = dict(enumerate(_TABLE.names))
        self.original_task = np.array(NOMINAL_TASK)
        self.task_dim = 4
        self.min_task = np.zeros(4)
        self.max_task = np.zeros(4)
        self.mean_task = np.zeros(4)
        self.stdev_task = np.zeros(4)
        self.reward_threshold = _TABLE.reward_threshold
        self.theta_threshold_radians = THETA_THRESHOLD_RADIANS
        self.x_threshold = X_THRESHOLD
        self.force_mag, self.tau = 10.0, 0.02
        self.polemass_length = 0.05     # frozen, as in the reference (:79 vs :157-166)
        self._buffers = None
        self._cfg_cache = None
        self._cfg_key = None
        self._tick = 0                  # step clock: Philox episode key; +1 per reset/step, +K per rollout
        self.seed_dr(self._seed)        # set_random_task / sample_tasks draw from the constructor seed's stream

    # ---- xi tables (random_cartpole.py:123-147) ----------------------------------------------------
    def get_search_bounds_mean(self, index):
        return _TABLE.search_bounds[index]

    def get_task_lower_bound(self, index):
        return _TABLE.lower_bounds[index]

    # ---- buffers ----------------------------------------------------------------------------------
    @property
    def torch_dtype(self):
        t = _device.torch()
        return t.float32 if self._dtype_name == "float32" else t.float64

    @property
    def device(self):
        return self._alloc()["device"]

    def _alloc(self):
        if self._buffers is not None:
            return self._buffers
        t = _device.torch()
        dev = _device.require_cuda(self._device_arg)
        _lib.load()
        n, ld = self.num_envs, _round_up(self.num_envs, 32)
        dt = self.torch_dtype
        b = dict(device=dev, ld=ld)
        b["state"] = t.zeros((4, ld), dtype=dt, device=dev)
        b["xi"] = t.tensor(NOMINAL_TASK, dtype=dt, device=dev).reshape(1, 4).repeat(ld, 1).contiguous()   # (ld, 4) rows
        b["elapsed"] = None if self.lean else t.zeros(ld, dtype=t.int32, device=dev)
        b["elapsed16"] = t.zeros(ld, dtype=t.int16, device=dev) if self.lean else None     # uint16 bit pattern
        b["episode"] = t.zeros(ld, dtype=t.int32, device=dev)       # uint32 bit pattern
        b["beyond"] = t.full((ld,), -1, dtype=t.int32, device=dev)
        b["reward"] = t.ones(ld, dtype=dt, device=dev)      # lean: never written, 1.0 on every step (:207-212)
        b["done"] = t.zeros(ld, dtype=t.uint8, device=dev)
        b["truncated"] = t.zeros(ld, dtype=t.uint8, device=dev)
        b["action"] = t.zeros(ld, dtype=t.uint8, device=dev)
        b["stats"] = t.tensor([0.0, 0.0, 0.0, math.inf, -math.inf, 0.0], dtype=t.float64, device=dev)
        b["obs"] = t.zeros((4, ld), dtype=dt, device=dev) if self.noisy else None
        tiles = -(-n // _lib.TILE_ENVS[self._dtype_name])
        b["progress"] = t.zeros(2 * tiles, dtype=t.int32, device=dev) if self.tile_ordering else None
        b["env"] = None
        self._buffers = b
        self._refresh_env_struct()
        return b

    def _refresh_env_struct(self):
        b = self._buffers
        env = _lib.CartpoleEnv()
        env.state, env.xi = b["state"].data_ptr(), b["xi"].data_ptr()
        env.elapsed = None if self.lean else b["elapsed"].data_ptr()
        env.elapsed16 = b["elapsed16"].data_ptr() if self.lean else None
        env.beyond = b["beyond"].data_ptr()
        env.progress = b["progress"].data_ptr() if b["progress"] is not None else None
        env.episode = b["episode"].data_ptr() if self.track_episodes else None
        env.n, env.ld = self.num_envs, b["ld"]
        env.env_id0, env.seed = self.env_id0, self._seed
        b["env"] = env
        b["step_plan"] = None          # holds byref(env): rebuild
        b["noise"] = None
        if self.noisy:
            b["noise"] = _lib.ObsNoise()
            b["noise"].obs, b["noise"].std = b["obs"].data_ptr(), math.sqrt(self.noise_level)

    def _entry(self, name):
        """(function name, leading arguments): the *_noisy entry points take the obs-noise block after env."""
        b = self._buffers
        if self.noisy:
            b["noise"].std = math.sqrt(self.noise_level)      # noise_level is a plain attribute in the reference
            return "renv_cartpole_%s_noisy_%s" % (name, self._suffix()), (ctypes.byref(b["env"]), ctypes.byref(b["noise"]))
        return "renv_cartpole_%s_%s" % (name, self._suffix()), (ctypes.byref(b["env"]),)

    def _suffix(self):
        return "f32" if self._dtype_name == "float32" else "f64"

    def _integrator(self):
        return _lib.EULER if self.kinematics_integrator == "euler" else _lib.SEMI_IMPLICIT   # :187-196

    def _on_distribution_change(self):
        self._cfg_cache = None

    def _active_dr_cfg(self):
        """DR config used on reset: only when dr_training is on and a distribution is loaded."""
        if not self.dr_training or self.sampling is None:
            return None
        # the reference reads min/max/mean/stdev_task at every draw, so in-place edits of those arrays must take
        # effect: the cached struct is keyed by their bytes (4 x 32 B, ~1 us)
        key = (self.sampling, self.min_task.tobytes(), self.max_task.tobytes(), self.mean_task.tobytes(),
               self.stdev_task.tobytes(),
               np.asarray(self.cov_task).tobytes() if self.sampling == "fullgaussian" else None)
        if self._cfg_cache is None or self._cfg_key != key:
            self._cfg_cache, self._cfg_key = self.dr_config(), key
        return ctypes.byref(self._cfg_cache)

    # ---- gym-style API ------------------------------------------------------------------------------
    def seed(self, seed=None):
        """Philox key for initial states and DR draws (random_cartpole.py:168-170 returns [seed])."""
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 63))
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.seed_dr(self._seed)
        self._tick = 0                  # the reference rebuilds np_random: seed(s); reset() repeats the same episode
        if self._buffers is not None:
            self._refresh_env_struct()
        return [seed]

    def reset(self, mask=None):
        """Reset every env (or those with mask[i] != 0).  Returns obs (N, 4)."""
        b = self._alloc()
        t = _device.torch()
        mask_ptr = None
        if mask is not None:
            mask = t.as_tensor(mask, device=b["device"]).to(t.uint8).contiguous()
            assert mask.shape == (self.num_envs,)
            mask_ptr = _device.ptr(mask)
        viol = self._violation_counter(b["device"])
        fn, head = self._entry("reset")
        with t.cuda.device(b["device"]):
            _lib.call(fn, *head, mask_ptr, self._tick,
                      self._active_dr_cfg(), _device.ptr(viol), _device.stream_ptr(b["device"]))
        self._tick += 1
        return self.obs

    @property
    def obs(self):
        """(N, 4) view of the last observation: the state itself, or the noisy copy for the Noisy variant."""
        b = self._alloc()
        return (b["obs"] if self.noisy else b["state"])[:, :self.num_envs].t()

    @property
    def state(self):
        """(N, 4) view of the true state (what the dynamics integrate; equals ``obs`` unless ``noisy``)."""
        return self._alloc()["state"][:, :self.num_envs].t()

    def _stage_actions(self, actions):
        b = self._buffers
        t = _device.torch()
        n = self.num_envs
        if isinstance(actions, t.Tensor) and actions.is_cuda and actions.dtype == t.uint8 and actions.is_contiguous() \
                and actions.data_ptr() % 16 == 0 and actions.device == b["device"] \
                and (tuple(actions.shape) == (n,) or actions is b["action"]):
            return actions              # zero copy; values outside {0, 1} are flagged by the step kernel itself
        src = actions if isinstance(actions, t.Tensor) else t.as_tensor(np.asarray(actions))
        if tuple(src.shape) != (n,):
            raise ValueError("actions must have shape (%d,), got %s" % (n, tuple(src.shape)))
        if src.dtype.is_floating_point:
            raise AssertionError("%r (%s) invalid" % (actions, type(actions)))   # Discrete(2) rejects floats
        if src.dtype == t.bool:         # a comparison result, e.g. (obs[:, 2] > 0): 0 / 1 by construction
            src = src.to(t.uint8)
        elif src.dtype != t.uint8:
            # a wider integer would be truncated to 8 bits by the staging copy: clamp so that anything outside
            # {0, 1} stays outside (negative -> 255, large -> 255) and is caught on the device
            src = src.to(b["device"], non_blocking=True)
            src = t.where((src < 0) | (src > 255), t.full_like(src, 255), src)
        b["action"][:n].copy_(src, non_blocking=True)
        return b["action"]

    def _step_plan(self):
        """Everything about a ``step`` launch that does not change between calls: the bound C function, the ctypes
        pointers of the persistent buffers and the views handed back.  Rebuilt when a layout-affecting attribute
        changes.  (A 2^20-env step is an 11 us kernel: per-call Python work has to stay below that.)"""
        b = self._buffers
        key = (self.noisy, self.track_truncated, self._dtype_name, self.lean)
        plan = b.get("step_plan")
        if plan is not None and plan["key"] == key:
            return plan
        t = _device.torch()
        n = self.num_envs
        fn, head = self._entry("step_lean" if self.lean else "step")
        info = {"TimeLimit.truncated": b["truncated"][:n].view(t.bool)} if self.track_truncated else {}
        viol = self._violation_counter(b["device"])
        plan = dict(key=key, fn=getattr(_lib.load(), fn), name=fn, head=head,
                    reward=_device.ptr(b["reward"]), done=_device.ptr(b["done"]),
                    truncated=_device.ptr(b["truncated"]) if self.track_truncated else None,
                    out=(self.obs, b["reward"][:n], b["done"][:n].view(t.bool)), info=info,
                    device_index=b["device"].index, viol=ctypes.c_void_p(viol.data_ptr()), viol_tensor=viol,
                    own_action=b["action"], own_action_ptr=_device.ptr(b["action"]))
        b["step_plan"] = plan
        return plan

    def step(self, actions):
        """One env-step for all N envs.  actions: (N,) integer tensor/array in {0, 1}.

        Returns (obs (N, 4), reward (N,), done (N,) bool, info) -- views of the env's own buffers, refreshed in place
        by every step (as the reference's obs aliases its state, random_cartpole.py:224).

        Nothing here synchronises with the GPU, so the two errors the reference raises from inside ``step`` /
        ``reset`` -- an action outside Discrete(2) (random_cartpole.py:173-174) and a gaussian DR draw that stayed
        below 0.1 three times (random_env.py:181-186) -- are detected by the kernels, counted on the device and raised
        by the next call that already synchronises: ``step_host_wait``, ``episode_stats``, ``state_dict``,
        ``set_random_task``, ``sample_tasks`` or an explicit ``check_dr_violations()``."""
        b = self._alloc()
        t = _device.torch()
        plan = self._step_plan()
        if actions is plan["own_action"] or actions is b.get("action_view"):   # sample_actions(): nothing to check or copy
            action_ptr = plan["own_action_ptr"]
        else:
            action_ptr = ctypes.c_void_p(self._stage_actions(actions).data_ptr())
        if self.noisy:
            b["noise"].std = math.sqrt(self.noise_level)
        idx = plan["device_index"]
        if self.lean:
            args = plan["head"] + (action_ptr, plan["done"], plan["truncated"], self._integrator(), self.max_episode_steps,
                                   self._tick, self._active_dr_cfg(), plan["viol"], ctypes.c_void_p(_device.raw_stream(idx)))
        else:
            args = plan["head"] + (action_ptr, plan["reward"], plan["done"], plan["truncated"],
                                   self._integrator(), self.max_episode_steps, int(self.auto_reset), self._tick,
                                   self._active_dr_cfg(), plan["viol"], ctypes.c_void_p(_device.raw_stream(idx)))
        if t.cuda.current_device() == idx:
            rc = plan["fn"](*args)
        else:
            with t.cuda.device(b["device"]):
                rc = plan["fn"](*args)
        if rc != _lib.OK:
            raise _lib.RenvError(plan["name"], rc, _lib.strerror(rc))
        self._tick += 1
        obs, reward, done = plan["out"]
        return obs, reward, done, dict(plan["info"])

    def _own_action_view(self):
        """(N,) view of the env's own action buffer -- one cached object, recognised by ``step`` without any checks."""
        b = self._buffers
        v = b.get("action_view")
        if v is None:
            v = b["action_view"] = b["action"][:self.num_envs]
        return v

    def sample_actions(self, out=None):
        """``action_space.sample()`` for every env: (N,) uint8 Bernoulli(1/2), Philox keyed by the step clock."""
        b = self._alloc()
        t = _device.torch()
        out = b["action"] if out is None else out
        with t.cuda.device(b["device"]):
            _lib.call("renv_random_actions_u8", _device.ptr(out), self.num_envs, self.env_id0, self._seed,
                      self._tick & 0xFFFFFFFF, _device.stream_ptr(b["device"]))
        return self._own_action_view() if out is b["action"] else out

    def rollout(self, w, b=0.0, num_steps=MAX_EPISODE_STEPS):
        """K fused env-steps under the in-kernel linear policy a = [w.s + b > 0] (auto-reset always on), or, with
        ``w=None``, under the random policy ``action_space.sample()`` (test_random_policy.py:26)."""
        buf = self._alloc()
        t = _device.torch()
        if self.lean:
            raise ValueError("rollout() keeps the int32 TimeLimit counter: build the env without lean=True")
        if w is None:         # random policy: the fused equivalent of K x step(sample_actions())
            if self.noisy:
                raise ValueError("the random policy ignores observations: use a noise-free env")
            w_arr = None
        else:
            w_arr = (ctypes.c_double * 4)(*[float(v) for v in w])
        viol = self._violation_counter(buf["device"])
        fn, head = self._entry("rollout")         # Noisy variant: the policy acts on the noisy observation
        with t.cuda.device(buf["device"]):
            _lib.call(fn, *head, w_arr, float(b),
                      int(num_steps), self._integrator(), self.max_episode_steps, self._tick, self._active_dr_cfg(),
                      _device.ptr(buf["stats"]), _device.ptr(viol), _device.stream_ptr(buf["device"]))
        self._tick += int(num_steps)

    @property
    def stats_tensor(self):
        """Device tensor [episodes, sum R, sum R^2, min R, max R, sum length] (float64)."""
        return self._alloc()["stats"]

    def reset_stats(self):
        t = _device.torch()
        self._alloc()["stats"].copy_(t.tensor([0.0, 0.0, 0.0, math.inf, -math.inf, 0.0], dtype=t.float64))

    def allgather_stats(self, group=None):
        """The per-iteration collective (distributed.allgather_stats) on this env's statistics; like every call that
        synchronises it also surfaces the device-side error flags."""
        from .distributed import allgather_stats
        combined, gathered = allgather_stats(self.stats_tensor, group)
        host = combined.cpu()                   # synchronises
        self.check_dr_violations()
        return host, gathered

    def episode_stats(self):
        from .distributed import summarize_stats
        stats = self.stats_tensor.cpu().numpy()         # synchronises: the deferred device-side errors surface here
        self.check_dr_violations()
        return summarize_stats(stats)

    def check_dr_violations(self):
        """Raise what the reference would have raised inside step()/reset() (see ``step``); ``validate_actions=False``
        silences the invalid-action assertion (the device flag is then cleared unseen)."""
        if not self.validate_actions and self._dr_violations is not None:
            self._dr_violations[_lib.COUNTER_BAD_ACTION] = 0
        RandomEnv.check_dr_violations(self)

    # ---- tasks --------------------------------------------------------------------------------------
    def get_task(self):
        return self._alloc()["xi"][:self.num_envs]

    def set_task(self, *task):
        """set_task(Tensor(N, 4)) for per-env xi, or set_task(g, m_c, m_p, l) to broadcast one task."""
        b = self._alloc()
        t = _device.torch()
        n = self.num_envs
        if len(task) == 1:
            xi = t.as_tensor(task[0], device=b["device"]).to(self.torch_dtype)
            if xi.shape == (4,):
                xi = xi.reshape(1, 4).expand(n, 4)
            if xi.shape != (n, 4):
                raise ValueError("set_task expects (N, 4) or 4 scalars")
            b["xi"][:n].copy_(xi)
        elif len(task) == 4:
            b["xi"][:n].copy_(t.tensor([float(v) for v in task], dtype=self.torch_dtype,
                                       device=b["device"]).reshape(1, 4).expand(n, 4))
        else:
            raise ValueError("set_task expects (N, 4) or 4 scalars")

    def set_random_task(self):
        """Resample xi of every env now (random_env.py:37-39)."""
        # sample i of the call is keyed by the GLOBAL env id, so a sharded env draws what the unsharded one would
        self.set_task(self.sample_tasks_tensor(self.num_envs, dtype=self.torch_dtype, device=self._alloc()["device"],
                                               sample_id0=self.env_id0))
        self.check_dr_violations()

    # ---- host-buffer entry point (end-to-end path: H2D actions, D2H results) -------------------------
    def host_buffers(self):
        """Pinned host staging buffers (numpy views): write ``action`` in place to skip one host memcpy."""
        b = self._alloc()
        t = _device.torch()
        n = self.num_envs
        h = b.get("host")
        if h is None:
            h = dict(action=t.empty(b["ld"], dtype=t.uint8).pin_memory(),
                     state=t.empty((4, b["ld"]), dtype=self.torch_dtype).pin_memory(),
                     reward=t.empty(b["ld"], dtype=self.torch_dtype).pin_memory(),
                     done=t.empty(b["ld"], dtype=t.uint8).pin_memory(),
                     truncated=t.empty(b["ld"], dtype=t.uint8).pin_memory(),
                     counters=t.zeros(_lib.NUM_COUNTERS, dtype=t.int64).pin_memory())
            # flag bits: [0] done, [1] truncated, ceil(ld / 32) words each (device twin in b["flag_bits"])
            words = (b["ld"] + 31) // 32
            h["flag_bits"] = t.zeros((2, words), dtype=t.int32).pin_memory()
            b["flag_bits"] = t.zeros((2, words), dtype=t.int32, device=b["device"])
            h["flag_bytes"] = h["flag_bits"].numpy().view(np.uint8)          # (2, 4 * words)
            h["np"] = dict(action=h["action"].numpy()[:n], obs=h["state"].numpy()[:, :n].T,
                           reward=h["reward"].numpy()[:n], done=h["done"].numpy()[:n].view(np.bool_),
                           truncated=h["truncated"].numpy()[:n].view(np.bool_))
            h["reward"].fill_(1)      # with auto-reset the reward is 1.0 on every step (:207-212): never copied
            h["stream"] = t.cuda.Stream(device=b["device"])
            h["event"] = t.cuda.Event()
            b["host"] = h
        return h["np"]

    def step_host_async(self, actions=None):
        """Enqueue one host-buffer step (H2D actions -> step kernel -> D2H obs/reward/done[/truncated]) on this
        env's private stream and return immediately; ``step_host_wait`` returns the numpy results.

        ``actions``: numpy/sequence of {0,1} copied into the pinned staging buffer, or None when the caller has
        already written ``host_buffers()['action']``.  Several envs can have steps in flight at once, which
        overlaps one env's device->host transfer with another's host->device transfer and kernel.
        """
        views = self.host_buffers()
        b = self._buffers
        h = b["host"]
        t = _device.torch()
        if actions is not None:
            acts = np.asarray(actions)
            if acts.dtype != np.uint8:
                if acts.dtype.kind not in "iu":
                    raise AssertionError("%r (%s) invalid" % (actions, type(actions)))   # Discrete(2) rejects floats / bools
                acts = np.clip(acts, -1, 2)     # the cast below wraps mod 256: keep invalid values invalid (255 / 2)
            np.copyto(views["action"], acts, casting="unsafe")
        stream = h["stream"]
        stream.wait_stream(t.cuda.current_stream(b["device"]))
        with t.cuda.stream(stream):
            b["action"].copy_(h["action"], non_blocking=True)
            self.step(b["action"])
            h["state"].copy_(b["obs"] if self.noisy else b["state"], non_blocking=True)
            if not self.auto_reset:        # only the steps-beyond-done rule (:213-222) ever yields 0.0
                h["reward"].copy_(b["reward"], non_blocking=True)
            if self.pack_host_flags:
                sp = _device.stream_ptr(b["device"])
                bits = b["flag_bits"]
                _lib.call("renv_pack_flags_u8", _device.ptr(b["done"]), _device.ptr(bits[0]), self.num_envs, sp)
                if self.track_truncated:
                    _lib.call("renv_pack_flags_u8", _device.ptr(b["truncated"]), _device.ptr(bits[1]), self.num_envs, sp)
                    h["flag_bits"].copy_(bits, non_blocking=True)
                else:
                    h["flag_bits"][0].copy_(bits[0], non_blocking=True)
            else:
                h["done"].copy_(b["done"], non_blocking=True)
                if self.track_truncated:
                    h["truncated"].copy_(b["truncated"], non_blocking=True)
            h["counters"].copy_(self._violation_counter(b["device"]), non_blocking=True)    # 16 bytes: the error flags
            h["event"].record(stream)

    def host_bytes_per_step(self):
        """(host->device, device->host) bytes one ``step_host`` moves over PCIe."""
        esize = 4 if self._dtype_name == "float32" else 8
        ld = self._alloc()["ld"]
        flag = 4 * ((ld + 31) // 32) if self.pack_host_flags else ld
        d2h = 4 * ld * esize + flag + (0 if self.auto_reset else ld * esize) + (flag if self.track_truncated else 0) \
            + 8 * _lib.NUM_COUNTERS
        return ld, d2h

    def step_host_wait(self):
        """Block until the last ``step_host_async`` finished -> (obs (N,4), reward, done, truncated) numpy views."""
        h = self._buffers["host"]
        h["event"].synchronize()
        c = h["counters"]
        if int(c[0]) | int(c[1]):          # the device flagged a gaussian failure or an invalid action (see ``step``)
            self.check_dr_violations()
        v = h["np"]
        if self.pack_host_flags:
            n = self.num_envs
            np.copyto(v["done"].view(np.uint8), np.unpackbits(h["flag_bytes"][0], count=n, bitorder="little"))
            if self.track_truncated:
                np.copyto(v["truncated"].view(np.uint8), np.unpackbits(h["flag_bytes"][1], count=n, bitorder="little"))
        return v["obs"], v["reward"], v["done"], v["truncated"]

    def step_host(self, actions):
        """``step`` with HOST buffers: numpy uint8 actions in, numpy (obs, reward, done, truncated) out."""
        self.step_host_async(actions)
        return self.step_host_wait()

    # ---- checkpoint / resume --------------------------------------------------------------------------
    def _tensor_keys(self):
        return ("state", "xi", "elapsed16" if self.lean else "elapsed", "episode", "beyond", "stats") + \
               (("obs",) if self.noisy else ())

    def state_dict(self):
        """Everything a resumed run needs to continue bit for bit: env buffers, the Philox key and step clock, the DR
        sampler stream (seed_dr seed, call index), the loaded distribution and the dr_training flag."""
        b = self._alloc()
        out = {k: b[k].clone() for k in self._tensor_keys()}     # the device->host free clone still orders after the kernels
        out.update(seed=self._seed, env_id0=self.env_id0, tick=self._tick, num_envs=self.num_envs,
                   dtype=self._dtype_name, lean=self.lean,
                   dr=dict(seed=self._dr_seed, calls=self._dr_calls, sampling=self.sampling, dr_training=self.dr_training,
                           min_task=self.min_task.copy(), max_task=self.max_task.copy(), mean_task=self.mean_task.copy(),
                           stdev_task=self.stdev_task.copy(),
                           cov_task=np.copy(self.cov_task) if getattr(self, "cov_task", None) is not None else None))
        self.check_dr_violations()          # synchronises; a checkpoint must not hide an error the reference raises
        return out

    def load_state_dict(self, sd):
        if sd["num_envs"] != self.num_envs or sd["dtype"] != self._dtype_name or sd.get("lean", False) != self.lean:
            raise ValueError("state_dict is for %d %s envs%s" % (sd["num_envs"], sd["dtype"], " (lean)" if sd.get("lean") else ""))
        b = self._alloc()
        for k in self._tensor_keys():
            b[k].copy_(sd[k])
        self._seed, self.env_id0, self._tick = sd["seed"], sd["env_id0"], sd["tick"]
        dr = sd.get("dr")
        if dr is not None:
            self._dr_seed, self._dr_calls = dr["seed"], dr["calls"]
            self.sampling, self.dr_training = dr["sampling"], dr["dr_training"]
            self.min_task[:], self.max_task[:] = dr["min_task"], dr["max_task"]
            self.mean_task[:], self.stdev_task[:] = dr["mean_task"], dr["stdev_task"]
            if dr["cov_task"] is not None:
                self.cov_task = np.copy(dr["cov_task"])
            self._on_distribution_change()
        self._refresh_env_struct()

    # ---- introspection used by tests -------------------------------------------------------------------
    @property
    def elapsed(self):
        b = self._alloc()
        if self.lean:
            t = _device.torch()
            return b["elapsed16"][:self.num_envs].to(t.int32) & 0xFFFF
        return b["elapsed"][:self.num_envs]

    @property
    def episode(self):
        return self._alloc()["episode"][:self.num_envs]

    @property
    def steps_beyond_done(self):
        return self._alloc()["beyond"][:self.num_envs]

    def set_state(self, state, elapsed=None):
        """Inject per-env states (N, 4) (parity tests: the reference's ``env.state = s0``)."""
        b = self._alloc()
        t = _device.torch()
        st = t.as_tensor(state, device=b["device"]).to(self.torch_dtype)
        b["state"][:, :self.num_envs].copy_(st.t())
        b["beyond"].fill_(-1)
        if elapsed is not None:
            el = t.as_tensor(elapsed, device=b["device"]).to(t.int32)
            if self.lean:
                b["elapsed16"][:self.num_envs].copy_(el.to(t.int16))
            else:
                b["elapsed"][:self.num_envs].copy_(el)
