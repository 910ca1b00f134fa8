/* renv.h -- C ABI of librenv_b200.so: the B200-native RandomCartPole-v0 / DR-sampler hot path.
 *
 * This is the drop-in boundary.  The reference (gabrieletiboni/random-envs) is pure Python and has
 * no FFI; each entry point below names the Python method(s) of the reference it replaces
 * (paths relative to the reference tree).  The shipped binding is ctypes
 * (random_envs_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - Plain C: pointers, sizes, PODs.  No C++/torch types, no exceptions, no global state.
 *   - Every data pointer is a DEVICE pointer owned by the caller; the library never allocates,
 *     frees or retains memory.  `renv_dr_cfg` / `renv_cartpole_env` are HOST structs read
 *     during the call only.
 *   - Work is enqueued on the caller's `cudaStream_t` (passed as void*); no host sync inside.
 *   - Return value: 0 = OK; < 0 = invalid argument (enum renv_status); > 0 = cudaError_t of the launch.
 *   - Alignment: `state`, `xi`, `reward`, `elapsed`, `out` 16 bytes; `action`, `done`, `truncated`,
 *     `mask` 4 bytes; `ld` a multiple of 4 (f32) or 2 (f64).  Violations return RENV_E_ALIGN.
 *   - Layout: `state` is structure-of-arrays (4, ld): rows x, x_dot, theta, theta_dot.  `xi` is one row per env,
 *     (n, 4) row-major: columns gravity, cart_mass, pole_mass, pole_length (order of
 *     random_envs/random_cartpole.py:104-107,149-155) -- the layout of get_task() / sample_tasks(n), and a reset
 *     rewrites one 32-byte sector instead of four.
 *   - Errors the reference raises from inside the computation are detected on the device and COUNTED in
 *     `counters` (device pointer to RENV_NUM_COUNTERS uint64, caller-zeroed, may be NULL; the entry points that can
 *     only produce the first kind call it `violations`):
 *       counters[0]  gaussian DR dims whose three draws were all < 0.1 (random_env.py:181-186: the reference raises
 *                    Exception('Not all samples were above > 0.1 after 2 attempts'); here the dim is set to 0.1),
 *       counters[1]  step launches x threads that saw an action outside {0, 1} (random_cartpole.py:173-174: the
 *                    reference asserts; here the env is pushed left and the flag is raised).
 *       counters[2]  step CTAs whose tile-ordering wait (renv_cartpole_env.progress) timed out: the progress words
 *                    were not in the state the protocol leaves them in (see there); the step still ran.
 *     The host reads them at its next synchronisation point and raises the reference's exception.
 *   - RNG: counter-based Philox4x32-10, key = seed, counter = (global env id, tick, purpose|slot).  `tick` is
 *     the caller's step clock (48 bits used): pass a value that grows by 1 per reset/step call and by K per
 *     K-step rollout (step k of a rollout uses tick + k).  An env starts at most one episode per tick, so
 *     (seed, env_id0 + i, tick) names the episode; results never depend on launch geometry or sharding,
 *     and a K-step rollout at tick t equals K single steps at ticks t .. t+K-1 bit for bit.
 */
#ifndef RENV_B200_H
#define RENV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RENV_ABI_VERSION 6
#define RENV_MAX_DIM 32          /* largest task_dim in the suite is 30 (jinja/random_humanoid.py) */
#define RENV_NUM_STATS 6         /* episodes, sum R, sum R^2, min R, max R, sum length */
#define RENV_TILE_ENVS_F32 1024  /* envs per step CTA: the granule of renv_cartpole_env.progress */
#define RENV_TILE_ENVS_F64 512
#define RENV_NUM_COUNTERS 3      /* device-detected error conditions, see `counters` below */

enum renv_status {
    RENV_OK = 0,
    RENV_E_NULL = -1,            /* required pointer is NULL */
    RENV_E_ALIGN = -2,           /* pointer or ld violates the alignment contract */
    RENV_E_SIZE = -3,            /* n <= 0, ld < n, K <= 0 ... */
    RENV_E_DIM = -4,             /* dim outside [1, RENV_MAX_DIM] (or != 4 for cartpole) */
    RENV_E_DRTYPE = -5,          /* unknown dr_type (random_env.py:90 'Unknown dr_type') */
    RENV_E_INTEGRATOR = -6,
    RENV_E_ARG = -7
};

/* random_envs/random_env.py:72-90 -- the string keys of set_dr_distribution. */
enum renv_dr_type {
    RENV_DR_NONE = 0,            /* sampling is None / dr_training False: xi is left alone on reset */
    RENV_DR_UNIFORM = 1,         /* random_env.py:150-151 */
    RENV_DR_TRUNCNORM = 2,       /* random_env.py:153-171 */
    RENV_DR_GAUSSIAN = 3,        /* random_env.py:173-190 */
    RENV_DR_FULLGAUSSIAN = 4     /* random_env.py:192-198 + denormalize_parameters :205-220 */
};

/* random_envs/random_cartpole.py:187-196: 'euler' vs anything else. */
enum renv_integrator { RENV_EULER = 0, RENV_SEMI_IMPLICIT = 1 };

/* Host-side image of RandomEnv's distribution state (random_env.py:102-121 de-interleaved):
 *   uniform:            a = min_task,  b = max_task
 *   truncnorm/gaussian: a = mean_task, b = stdev_task
 *   lb[i] = get_task_lower_bound(i) (used by truncnorm only; gaussian's floor is the literal 0.1).
 *   fullgaussian:       a = mean_task in the normalised [0, 4] space, b / lb = get_task_search_bounds() lo / hi,
 *                       factor = row-major dim x dim matrix F with F F^T = cov_task (e.g. its Cholesky factor):
 *                       xi = denormalize(clip(a + F z, 0, 4)), z ~ N(0, I). */
typedef struct renv_dr_cfg {
    int32_t dr_type;
    int32_t dim;
    double a[RENV_MAX_DIM];
    double b[RENV_MAX_DIM];
    double lb[RENV_MAX_DIM];
    double factor[RENV_MAX_DIM * RENV_MAX_DIM];
} renv_dr_cfg;

/* One shard of cart-pole envs resident in HBM (device pointers, element type T = float | double). */
typedef struct renv_cartpole_env {
    void *state;                 /* T (4, ld)   RandomCartPoleEnv.state            :176,198 */
    void *xi;                    /* T (n, 4)    gravity, cart_mass, pole_mass, pole_length :157-166 */
    int32_t *elapsed;            /* (n) TimeLimit._elapsed_steps (gym 0.21)                 */
    uint32_t *episode;           /* (n) episodes started per env (statistics only; may be NULL)   */
    int32_t *beyond;             /* (n) steps_beyond_done, -1 == None (:207-222); may be NULL when auto_reset */
    uint16_t *elapsed16;         /* (n) the same TimeLimit counter as uint16 for renv_cartpole_step_lean_f32 (16-byte
                                    aligned); NULL otherwise.  reset zeroes whichever of elapsed / elapsed16 is set;
                                    `elapsed` may be NULL only for an env that is stepped through the lean entry. */
    uint32_t *progress;          /* step-to-step ordering at TILE granularity, or NULL for plain stream order.
                                    2 * ceil(n / RENV_TILE_ENVS_F32 | _F64) uint32, 8-byte aligned, zero-initialised ONCE
                                    by the caller and afterwards touched only by the step entry points: word 2b counts
                                    the step CTAs that have started on tile b, word 2b+1 those that have finished.  A
                                    step CTA runs as soon as ITS tile's previous step has finished instead of waiting
                                    for the whole previous grid, so consecutive step launches of one stream overlap
                                    (also under CUDA-graph replay).  All step launches that share a progress array must
                                    be issued in one stream (or otherwise ordered), as for any in-place update. */
    int64_t n;                   /* envs in this shard */
    int64_t ld;                  /* row stride of state in elements, >= n */
    uint64_t env_id0;            /* global id of env 0 of this shard (rank * n under contiguous sharding) */
    uint64_t seed;               /* Philox key */
} renv_cartpole_env;

int renv_abi_version(void);
const char *renv_strerror(int code);

/* RandomEnv.sample_tasks(n) -> (n, dim) row-major  (random_env.py:145-203).
 * Sample i uses Philox id = sample_id0 + i and episode field = call.  `violations` (may be NULL) is the
 * RENV_NUM_COUNTERS-element counter array of the Conventions: [0] counts gaussian dims whose three draws were all
 * < 0.1 -- the host raises the reference's Exception('Not all samples were above > 0.1 after 2 attempts') when it
 * is non-zero.
 * dr_type fullgaussian with 17 <= dim <= 32 in fp32 runs its (n x 32) . (32 x 32) contraction on the tensor cores
 * (tcgen05.mma kind::tf32, split into head and tail: fp32-grade products, fp32 accumulation); smaller dims and fp64
 * use FMA chains.  Both draw the same normals; the results agree to 4e-6 of the search-bound width. */
int renv_dr_sample_f32(float *out, int64_t n, const renv_dr_cfg *cfg, uint64_t seed, uint64_t sample_id0,
                       uint32_t call, unsigned long long *violations, void *stream);
int renv_dr_sample_f64(double *out, int64_t n, const renv_dr_cfg *cfg, uint64_t seed, uint64_t sample_id0,
                       uint32_t call, unsigned long long *violations, void *stream);

/* RandomCartPoleEnv.reset (random_cartpole.py:226-229) for every env with mask[i] != 0 (mask NULL = all),
 * preceded by RandomEnv.set_random_task (random_env.py:37-39) when dr != NULL and dr->dr_type != NONE
 * (the README.md:9 / MuJoCo-env behaviour; CartPole's own reset forgets it).  Also zeroes elapsed,
 * sets beyond = -1 and counts the new episode. */
int renv_cartpole_reset_f32(const renv_cartpole_env *env, const uint8_t *mask, uint64_t tick, const renv_dr_cfg *dr,
                            unsigned long long *violations, void *stream);
int renv_cartpole_reset_f64(const renv_cartpole_env *env, const uint8_t *mask, uint64_t tick, const renv_dr_cfg *dr,
                            unsigned long long *violations, void *stream);

/* RandomCartPoleEnv.step (random_cartpole.py:172-224) for all n envs, fused with
 *   TimeLimit.step (gym 0.21; max_steps <= 0 disables)  and
 *   SyncVectorEnv auto-reset (gym 0.21; auto_reset != 0) incl. the DR resample above.
 * action (n) in {0,1}; reward (n) T; done (n) u8; truncated (n) u8 or NULL.
 * With auto_reset the state written back for a finished env is its reset state (the obs gym returns), and
 * `reward` may be NULL: the reward is then 1.0 for every env on every step (:207-212) and is not written.
 * `counters`: RENV_NUM_COUNTERS uint64 (see Conventions). */
int renv_cartpole_step_f32(const renv_cartpole_env *env, const uint8_t *action, float *reward, uint8_t *done,
                           uint8_t *truncated, int integrator, int max_steps, int auto_reset, uint64_t tick,
                           const renv_dr_cfg *dr, unsigned long long *counters, void *stream);
int renv_cartpole_step_f64(const renv_cartpole_env *env, const uint8_t *action, double *reward, uint8_t *done,
                           uint8_t *truncated, int integrator, int max_steps, int auto_reset, uint64_t tick,
                           const renv_dr_cfg *dr, unsigned long long *counters, void *stream);

/* The LEAN auto-reset step: the same env-step with 54 instead of 62 bytes of HBM traffic per env (fp32).  The
 * TimeLimit counter is env->elapsed16 (uint16: max_steps <= 65535) and no reward is written (identically 1.0 under
 * auto-reset, :207-212).  state, xi, done and truncated are bit-identical to renv_cartpole_step_f32 with
 * auto_reset = 1. */
int renv_cartpole_step_lean_f32(const renv_cartpole_env *env, const uint8_t *action, uint8_t *done, uint8_t *truncated,
                                int integrator, int max_steps, uint64_t tick, const renv_dr_cfg *dr,
                                unsigned long long *counters, void *stream);

/* Observation noise of the suite's "Noisy" env variants (jinja/random_hopper.py:28,107-108, random_half_cheetah.py,
 * random_walker2d.py:139-140, random_humanoid.py:193-204): every observation -- after a step and after a reset --
 * is  obs = state + sqrt(noise_level) * N(0, I);  the state itself is not perturbed.  The *_noisy entry points are
 * the plain ones plus this block: `obs` is a DEVICE buffer T (4, ld), 16-byte aligned, that receives the noisy
 * observation (with auto_reset a finished env's row holds the noisy observation of its reset state); `std` =
 * sqrt(noise_level) >= 0.  Normals are Philox draws keyed (seed, env id, tick, purpose 4). */
typedef struct renv_obs_noise {
    void *obs;                   /* T (4, ld) */
    double std;                  /* sqrt(noise_level) */
} renv_obs_noise;

int renv_cartpole_reset_noisy_f32(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *mask,
                                  uint64_t tick, const renv_dr_cfg *dr, unsigned long long *violations, void *stream);
int renv_cartpole_reset_noisy_f64(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *mask,
                                  uint64_t tick, const renv_dr_cfg *dr, unsigned long long *violations, void *stream);
int renv_cartpole_step_noisy_f32(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *action,
                                 float *reward, uint8_t *done, uint8_t *truncated, int integrator, int max_steps,
                                 int auto_reset, uint64_t tick, const renv_dr_cfg *dr, unsigned long long *counters,
                                 void *stream);
int renv_cartpole_step_noisy_f64(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *action,
                                 double *reward, uint8_t *done, uint8_t *truncated, int integrator, int max_steps,
                                 int auto_reset, uint64_t tick, const renv_dr_cfg *dr, unsigned long long *counters,
                                 void *stream);

/* K fused steps with the linear policy a = [w.s + b > 0] evaluated in-kernel, auto-reset always on.
 * State, xi and counters stay in registers for the K steps (1 <= K <= 2^30).  w == NULL selects the RANDOM policy
 * of the reference's demo loop (test_random_policy.py:26): at tick t env e takes the action bit renv_random_actions_u8
 * would give it for step = t, so the launch equals K x { renv_random_actions_u8; renv_cartpole_step } bit for bit.  stats (device, RENV_NUM_STATS doubles,
 * caller-initialised to {0,0,0,+inf,-inf,0}) is ACCUMULATED with the finished episodes' returns. */
int renv_cartpole_rollout_f32(const renv_cartpole_env *env, const double w[4], double b, int K, int integrator,
                              int max_steps, uint64_t tick, const renv_dr_cfg *dr, double *stats,
                              unsigned long long *violations, void *stream);
int renv_cartpole_rollout_f64(const renv_cartpole_env *env, const double w[4], double b, int K, int integrator,
                              int max_steps, uint64_t tick, const renv_dr_cfg *dr, double *stats,
                              unsigned long long *violations, void *stream);

/* The same with the Noisy variant's observation model: the policy acts on obs = state + std * N(0, I) -- for the
 * first step the content of noise->obs (what the last reset/step left there), afterwards the observation each step
 * or reset would have returned -- and noise->obs is left as K calls of step_noisy would leave it. */
int renv_cartpole_rollout_noisy_f32(const renv_cartpole_env *env, const renv_obs_noise *noise, const double w[4], double b,
                                    int K, int integrator, int max_steps, uint64_t tick, const renv_dr_cfg *dr,
                                    double *stats, unsigned long long *violations, void *stream);
int renv_cartpole_rollout_noisy_f64(const renv_cartpole_env *env, const renv_obs_noise *noise, const double w[4], double b,
                                    int K, int integrator, int max_steps, uint64_t tick, const renv_dr_cfg *dr,
                                    double *stats, unsigned long long *violations, void *stream);

/* The SCALAR drop-in env (gym.make('RandomCartPole-v0'): RandomCartPoleEnv.step / reset, random_cartpole.py:172-229,
 * one env, one host call per step) served by a resident one-warp kernel instead of one launch + synchronise per call.
 * `ctrl` is PINNED, DEVICE-MAPPED HOST memory (zero-initialised once): the host rings a request in by writing
 * `request = (seq << 8) | op` (seq grows by 1 per request, 24 bits; op 0 / 1 = step(action); 2 = reset, 3 = set state,
 * 4 = set xi, 5 = configure, 6 = exit, each reading the `arg` / `arg_u64` fields that the host wrote BEFORE the request
 * word -- see csrc/renv_scalar_server.cuh for the field meanings) and spins on `ack == seq`; the results are in
 * state / obs / xi / reward / done / beyond / violations (or, for a step, already in `next[action]`, see below).  `save` is RENV_SCALAR_SAVE_BYTES of zero-initialised DEVICE
 * memory that carries the env between kernel instances: the kernel is a lease -- after `lease_ns` without a request it
 * stores its registers there, sets ctrl->exited = lease_id and exits; the caller launches the next instance (lease_id
 * + 1) when it has a request and sees exited == the id it launched last.  One instance at a time per ctrl block. */
#define RENV_SCALAR_SAVE_BYTES 512
typedef struct renv_scalar_ctrl {
    uint32_t request;            /* host -> device */
    uint32_t pad0[15];
    double arg[16];              /* host -> device: arguments of ops >= 2 */
    uint64_t arg_u64[8];
    double state[4];             /* device -> host: x, x_dot, theta, theta_dot after the request */
    double obs[4];               /*   Noisy variant: state + sqrt(noise_level) N(0, I) */
    double xi[4];                /*   gravity, cart_mass, pole_mass, pole_length (written by ops >= 2) */
    double reward;
    int32_t done;
    int32_t beyond;              /*   steps_beyond_done, -1 == None */
    uint32_t violations;         /*   gaussian DR dims that failed three times (reset with resample) */
    uint32_t pad1;
    uint32_t ack;                /* device -> host: seq of the last request served */
    uint32_t exited;             /* device -> host: lease id of the instance that has exited */
    uint32_t pad2[16];
    /* Look-ahead (device -> host; enabled by arg_u64[5] != 0 of a configure request): after serving request `next_seq`
     * the kernel also publishes what step(0) and step(1) would return from the state it is in now.  A host that finds
     * next_seq == the seq of its last request takes next[action] as the result of its next step AT ONCE and rings that
     * step in as op 8 + action WITHOUT waiting (ops 8 / 9 are not acknowledged and write no result block; at most one
     * request is outstanding: before it rings again the host waits for next_seq == that step's seq), which takes the
     * PCIe round trip off the step's critical path.  The kernel counts the step clock itself for the Noisy variant's
     * observations (tick of the last reset + 1 per step, or arg_u64[4] of a configure request). */
    struct renv_scalar_outcome {
        double state[4];
        double obs[4];           /*   Noisy variant */
        double reward;
        int32_t done;
        int32_t beyond;
        uint32_t pad[12];
    } next[2];
    uint32_t next_seq;           /* release-stored after next[0..1] */
    uint32_t pad3[15];
} renv_scalar_ctrl;
int renv_cartpole_scalar_serve(renv_scalar_ctrl *ctrl, void *save, uint32_t lease_id, uint64_t lease_ns, void *stream);

/* action_space.sample() for n envs (test_random_policy.py:26): env e at clock `step` takes bit (step & 127) of its own
 * Philox block (e, step >> 7), purpose 2 -- one block holds an env's Bernoulli(1/2) actions for 128 consecutive steps. */
int renv_random_actions_u8(uint8_t *action, int64_t n, uint64_t env_id0, uint64_t seed, uint32_t step,
                           void *stream);

/* Host-path helper (no counterpart in the reference, whose envs live in host memory): packs n byte flags (0 / non-zero,
 * the `done` / `truncated` vectors of renv_cartpole_step_*) into bits -- bit (i & 7) of byte i >> 3, numpy's
 * bitorder "little" -- so that they cross PCIe as n / 8 bytes.  `bits` holds ceil(n / 32) 32-bit words; both pointers
 * 4-byte aligned (16-byte aligned flags take the vector path). */
int renv_pack_flags_u8(const uint8_t *flags, uint32_t *bits, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RENV_B200_H */
