// C ABI of librenv_b200.so (declared in include/renv.h): argument validation + kernel launches.
// Stateless and re-entrant: nothing is allocated, cached or synchronised here.
#include "../../include/renv.h"
#include "renv_kernels.cuh"
#include <cmath>
#include "renv_rollout_pair.cuh"
#include "renv_fullgauss_tc.cuh"
#include "renv_scalar_server.cuh"
#include <stdlib.h>

#ifndef RENV_ROLLOUT_F32_PAIR
#define RENV_ROLLOUT_F32_PAIR 1     // 0: one env per thread (scalar FFMA) for A/B timing
#endif

using namespace renv;

namespace {

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// Host-side conversion of the 4-dim DR configuration to the kernels' element type: the same casts, the same
// rounded-once `hi - lo` and the exact 2^-24 folding the device code used to do per reset.
template <typename T> int to_cfg4(const renv_dr_cfg *dr, DrCfg4<T> *out)
{
    out->dr_type = kDrNone;
    out->dim = 4;
    for (int k = 0; k < 4; ++k) { out->a[k] = T(0); out->b[k] = T(0); out->floor[k] = T(0); }
    for (int k = 0; k < 16; ++k) out->factor[k] = T(0);
    if (dr == nullptr || dr->dr_type == RENV_DR_NONE) return RENV_OK;
    if (dr->dr_type < RENV_DR_NONE || dr->dr_type > RENV_DR_FULLGAUSSIAN) return RENV_E_DRTYPE;
    if (dr->dim != 4) return RENV_E_DIM;
    out->dr_type = dr->dr_type;
    for (int k = 0; k < 4; ++k) {
        const T a = (T)dr->a[k], b = (T)dr->b[k];
        out->a[k] = a;
        if (dr->dr_type == RENV_DR_UNIFORM) {
            const T width = b - a;
            out->b[k] = sizeof(T) == 4 ? (T)(width * (T)(1.0 / 16777216.0)) : width;
            out->floor[k] = T(0);
        } else if (dr->dr_type == RENV_DR_FULLGAUSSIAN) {
            out->b[k] = b;                       // search-bound lo
            out->floor[k] = (T)dr->lb[k];        // search-bound hi
        } else {
            out->b[k] = b;
            out->floor[k] = dr->dr_type == RENV_DR_TRUNCNORM ? (T)dr->lb[k] : (T)0.1;
        }
    }
    if (dr->dr_type == RENV_DR_FULLGAUSSIAN)
        for (int k = 0; k < 16; ++k) out->factor[k] = (T)dr->factor[k];
    return RENV_OK;
}

enum ElapsedKind { kNeedElapsed32, kNeedElapsed16, kNeedEitherElapsed };
template <typename T>
int check_env(const renv_cartpole_env *env, bool need_beyond, EnvPtrs<T> *out, ElapsedKind elapsed = kNeedElapsed32)
{
    if (env == nullptr || env->state == nullptr || env->xi == nullptr) return RENV_E_NULL;
    if (elapsed == kNeedElapsed32 && env->elapsed == nullptr) return RENV_E_NULL;
    if (elapsed == kNeedElapsed16 && env->elapsed16 == nullptr) return RENV_E_NULL;
    if (elapsed == kNeedEitherElapsed && env->elapsed == nullptr && env->elapsed16 == nullptr) return RENV_E_NULL;
    if (need_beyond && env->beyond == nullptr) return RENV_E_NULL;
    if (env->n <= 0 || env->ld < env->n) return RENV_E_SIZE;
    constexpr int V = VecTraits<T>::V;
    if (env->ld % V != 0) return RENV_E_ALIGN;
    if (!aligned(env->state, 16) || !aligned(env->xi, 16) || !aligned(env->elapsed, 16) || !aligned(env->elapsed16, 16) ||
        (env->episode && !aligned(env->episode, 4)) || !aligned(env->progress, 8) || (env->beyond && !aligned(env->beyond, 4)))
        return RENV_E_ALIGN;
    out->state = static_cast<T *>(env->state);
    out->xi = static_cast<T *>(env->xi);
    out->elapsed = env->elapsed;
    out->elapsed16 = env->elapsed16;
    out->progress = env->progress;
    out->episode = env->episode;
    out->beyond = env->beyond;
    out->n = env->n;
    out->ld = env->ld;
    out->env_id0 = env->env_id0;
    out->seed = env->seed;
    out->obs = nullptr;
    out->noise_std = T(0);
    return RENV_OK;
}

// "Noisy" variants: attach the observation buffer.  noise == nullptr: plain env (obs is the state itself).
template <typename T> int attach_noise(const renv_obs_noise *noise, EnvPtrs<T> *env)
{
    if (noise == nullptr) return RENV_OK;
    if (noise->obs == nullptr) return RENV_E_NULL;
    if (!aligned(noise->obs, 16)) return RENV_E_ALIGN;
    if (!(noise->std >= 0.0)) return RENV_E_ARG;
    env->obs = static_cast<T *>(noise->obs);
    env->noise_std = (T)noise->std;
    return RENV_OK;
}

inline int launch_status() { return (int)cudaGetLastError(); }

#ifndef RENV_STEP_PDL
#define RENV_STEP_PDL 1
#endif
// Launch with programmatic stream serialization: the kernel may start while its predecessor in the stream drains and
// synchronises itself with griddepcontrol.wait (see cartpole_step_kernel).  Captured into CUDA graphs as a
// programmatic dependency edge.
template <typename Args>
int launch_pdl(void (*kernel)(const Args), unsigned blocks, unsigned threads, cudaStream_t stream, const Args &args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = RENV_STEP_PDL ? 1 : 0;
    const cudaError_t rc = cudaLaunchKernelEx(&cfg, kernel, args);
    return rc != cudaSuccess ? (int)rc : launch_status();
}

// fp32: the packed-transform kernel; fp64: the generic one
template <int kType, int kStore>
void launch_sampler(float *out, int64_t n, const DrCfgPrepared<float> &c, uint64_t seed, uint64_t sample_id0, uint32_t call,
                    unsigned long long *violations, int items, unsigned blocks, cudaStream_t st)
{
    const PhiloxKeys ks = philox_keys((uint32_t)seed, (uint32_t)(seed >> 32));
    dr_sample_f32_kernel<kType, kStore><<<blocks, kSampleThreads, 0, st>>>(out, n, c, seed, ks, sample_id0, call, violations, items);
}
template <int kType, int kStore>
void launch_sampler(double *out, int64_t n, const DrCfgPrepared<double> &c, uint64_t seed, uint64_t sample_id0, uint32_t call,
                    unsigned long long *violations, int items, unsigned blocks, cudaStream_t st)
{
    if (kType == kDrUniform) {
        // the fast loop pre-scales the widths by 2^-53: exact unless a width is subnormal-small or not finite
        bool plain = true;
        for (int k = 0; k < c.dim; ++k) plain = plain && (c.b[k] == 0.0 || (std::fabs(c.b[k]) > 1e-250 && std::isfinite(c.b[k])));
        if (plain) {
            const PhiloxKeys ks = philox_keys((uint32_t)seed, (uint32_t)(seed >> 32));
            dr_sample_f64_uniform_kernel<kStore><<<blocks, kSampleThreads, 0, st>>>(out, n, c, ks, sample_id0, call, items);
            return;
        }
    }
    dr_sample_kernel<double, kType, kStore><<<blocks, kSampleThreads, 0, st>>>(out, n, c, seed, sample_id0, call, violations, items);
}

// fullgaussian, fp32, 17 <= dim <= 32: the tcgen05 kernel (renv_fullgauss_tc.cuh), one wave of persistent CTAs.
// Returns false when the tensor path does not apply (fp64, or switched off for A/B tests) and the caller falls through
// to the CUDA-core kernel.
bool launch_fullgaussian_tc(double *, int64_t, const FullGaussCfg<double> &, uint64_t, uint64_t, uint32_t,
                            unsigned long long *, cudaStream_t, int *) { return false; }
bool launch_fullgaussian_tc(float *out, int64_t n, const FullGaussCfg<float> &g, uint64_t seed, uint64_t sample_id0,
                            uint32_t call, unsigned long long *counters, cudaStream_t st, int *rc)
{
    const char *knob = getenv("RENV_FULLGAUSS_TENSOR");           // "0": CUDA-core kernel (tests compare the two)
    if (knob && knob[0] == '0') return false;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { *rc = (int)e; return true; }
    const int64_t tiles = (n + kFgTile - 1) / kFgTile;
    const int64_t grid = tiles < (int64_t)sms * RENV_FG_TC_CTAS ? tiles : (int64_t)sms * RENV_FG_TC_CTAS;
    FullGaussCfg<float> gt = g;
    for (int k = 0; k < 32; ++k) {      // the kernel's epilogue takes the WIDTH hi - lo (rounded once, as denormalize does) and mean / 4 (exact)
        gt.hi[k] = g.hi[k] - g.lo[k];
        gt.mean[k] = 0.25f * g.mean[k];
    }
    const PhiloxKeys ks = philox_keys((uint32_t)seed, (uint32_t)(seed >> 32));
    dr_sample_fullgaussian_tc_kernel<<<(unsigned)grid, kFgThreads, kFgPadSmem, st>>>(out, n, gt, ks, sample_id0, call, counters);
    *rc = launch_status();
    return true;
}

template <typename T>
int dr_sample(T *out, int64_t n, const renv_dr_cfg *cfg, uint64_t seed, uint64_t sample_id0, uint32_t call,
              unsigned long long *violations, void *stream)
{
    int rc_tc = RENV_OK;
    if (out == nullptr || cfg == nullptr) return RENV_E_NULL;
    if (n <= 0) return RENV_E_SIZE;
    if (cfg->dim < 1 || cfg->dim > RENV_MAX_DIM) return RENV_E_DIM;
    if (cfg->dr_type < RENV_DR_UNIFORM || cfg->dr_type > RENV_DR_FULLGAUSSIAN) return RENV_E_DRTYPE;
    if (!aligned(out, 16)) return RENV_E_ALIGN;
    if (cfg->dr_type == RENV_DR_FULLGAUSSIAN) {
        FullGaussCfg<T> g;
        g.dim = cfg->dim;
        for (int k = 0; k < 32; ++k) {
            const bool ok = k < cfg->dim;
            g.mean[k] = ok ? (T)cfg->a[k] : T(0); g.lo[k] = ok ? (T)cfg->b[k] : T(0); g.hi[k] = ok ? (T)cfg->lb[k] : T(0);
        }
        for (int k = 0; k < 32; ++k)                 // transposed, zero-padded: ft[k][d] = F[d][k]
            for (int d = 0; d < 32; ++d)
                g.ft[k * 32 + d] = (k < cfg->dim && d < cfg->dim) ? (T)cfg->factor[d * cfg->dim + k] : T(0);
        const int64_t gblocks = (n + kFullGaussThreads - 1) / kFullGaussThreads;
        if (gblocks > 0x7fffffffLL) return RENV_E_SIZE;
        const cudaStream_t gst = static_cast<cudaStream_t>(stream);
        if (cfg->dim > 16 && launch_fullgaussian_tc(out, n, g, seed, sample_id0, call, violations, gst, &rc_tc)) return rc_tc;
        if (cfg->dim <= 4) dr_sample_fullgaussian_kernel<T, 4><<<(unsigned)gblocks, kFullGaussThreads, 0, gst>>>(out, n, g, seed, sample_id0, call);
        else if (cfg->dim <= 8) dr_sample_fullgaussian_kernel<T, 8><<<(unsigned)gblocks, kFullGaussThreads, 0, gst>>>(out, n, g, seed, sample_id0, call);
        else if (cfg->dim <= 16) dr_sample_fullgaussian_kernel<T, 16><<<(unsigned)gblocks, kFullGaussThreads, 0, gst>>>(out, n, g, seed, sample_id0, call);
        else dr_sample_fullgaussian_kernel<T, 32><<<(unsigned)gblocks, kFullGaussThreads, 0, gst>>>(out, n, g, seed, sample_id0, call);
        return launch_status();
    }
    // host-side image of renv_dr.cuh load_dim_block: same conversions, same order, done once per launch
    DrCfgPrepared<T> c;
    c.dr_type = cfg->dr_type;
    c.dim = cfg->dim;
    for (int k = 0; k < 32; ++k) {
        const bool ok = k < cfg->dim;
        const T a = ok ? (T)cfg->a[k] : T(0), b = ok ? (T)cfg->b[k] : T(0);
        c.a[k] = a;
        if (cfg->dr_type == RENV_DR_UNIFORM) {
            const T width = b - a;                                          // rounded once, as numpy's `high - low`
            c.b[k] = sizeof(T) == 4 ? (T)(width * (T)(1.0 / 16777216.0)) : (T)width;   // fp32: 2^-24 folded in (exact)
        } else {
            c.b[k] = b;
        }
        c.floor[k] = !ok ? T(0) : cfg->dr_type == RENV_DR_TRUNCNORM ? (T)cfg->lb[k] : (T)0.1;
    }
    const int items = sampler_items<T>(n, cfg->dim);
    const int kTile = tile_samples<T>(cfg->dim, items);
    const int64_t blocks = (n + kTile - 1) / kTile;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    // store width follows the row alignment: whole 128-bit blocks, float pairs (e.g. the 30-dim humanoid), scalars
    constexpr int P = Pack<T>::kPerBlock;
    const int store = cfg->dim % P == 0 ? 0 : (sizeof(T) == 4 && cfg->dim % 2 == 0 ? 1 : 2);
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
#define RENV_LAUNCH_SAMPLER(TYPE, STORE) launch_sampler<TYPE, STORE>(out, n, c, seed, sample_id0, call, violations, items, (unsigned)blocks, st)
#define RENV_LAUNCH_SAMPLER_TYPE(TYPE) \
    do { if (store == 0) RENV_LAUNCH_SAMPLER(TYPE, 0); else if (store == 1) RENV_LAUNCH_SAMPLER(TYPE, 1); \
         else RENV_LAUNCH_SAMPLER(TYPE, 2); } while (0)
    if (cfg->dr_type == RENV_DR_UNIFORM) RENV_LAUNCH_SAMPLER_TYPE(kDrUniform);
    else if (cfg->dr_type == RENV_DR_TRUNCNORM) RENV_LAUNCH_SAMPLER_TYPE(kDrTruncnorm);
    else RENV_LAUNCH_SAMPLER_TYPE(kDrGaussian);
#undef RENV_LAUNCH_SAMPLER_TYPE
#undef RENV_LAUNCH_SAMPLER
    return launch_status();
}

template <typename T>
int cartpole_reset(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *mask, uint64_t tick,
                   const renv_dr_cfg *dr, unsigned long long *violations, void *stream)
{
    ResetArgs<T> a;
    int rc = check_env<T>(env, false, &a.env, kNeedEitherElapsed);
    if (rc) return rc;
    rc = attach_noise<T>(noise, &a.env);
    if (rc) return rc;
    rc = to_cfg4<T>(dr, &a.dr);
    if (rc) return rc;
    a.mask = mask;
    a.tick = tick;
    a.violations = violations;
    const int64_t blocks = (env->n + 255) / 256;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    cartpole_reset_kernel<T><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return launch_status();
}

template <typename T, bool kLean>
int cartpole_step(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *action, T *reward,
                  uint8_t *done, uint8_t *truncated, int integrator, int max_steps, int auto_reset, uint64_t tick,
                  const renv_dr_cfg *dr, unsigned long long *counters, void *stream)
{
    StepArgs<T> a;
    int rc = check_env<T>(env, !auto_reset, &a.env, kLean ? kNeedElapsed16 : kNeedElapsed32);
    if (rc) return rc;
    rc = attach_noise<T>(noise, &a.env);
    if (rc) return rc;
    if (action == nullptr || done == nullptr) return RENV_E_NULL;
    if (reward == nullptr && !auto_reset) return RENV_E_NULL;      // only the auto-reset reward is the constant 1.0
    if (!aligned(action, 4) || !aligned(reward, 16) || !aligned(done, 4) || (truncated && !aligned(truncated, 4)))
        return RENV_E_ALIGN;
    if (counters && !aligned(counters, 8)) return RENV_E_ALIGN;
    if (integrator != RENV_EULER && integrator != RENV_SEMI_IMPLICIT) return RENV_E_INTEGRATOR;
    if (kLean && max_steps > 65535) return RENV_E_SIZE;
    rc = to_cfg4<T>(auto_reset ? dr : nullptr, &a.dr);
    if (rc) return rc;
    a.action = action; a.reward = reward; a.done = done; a.truncated = truncated;
    a.euler = integrator == RENV_EULER;
    a.max_steps = max_steps;
    a.tick = tick;
    a.counters = counters;
    constexpr int64_t per_block = (int64_t)kStepThreads * VecTraits<T>::V;
    const int64_t blocks = (env->n + per_block - 1) / per_block;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    const cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool noisy = a.env.obs != nullptr;
    if (kLean) return launch_pdl(cartpole_step_kernel<T, true, false, true>, (unsigned)blocks, kStepThreads, st, a);
    if (auto_reset && !noisy) return launch_pdl(cartpole_step_kernel<T, true, false>, (unsigned)blocks, kStepThreads, st, a);
    if (auto_reset) return launch_pdl(cartpole_step_kernel<T, true, true>, (unsigned)blocks, kStepThreads, st, a);
    if (!noisy) return launch_pdl(cartpole_step_kernel<T, false, false>, (unsigned)blocks, kStepThreads, st, a);
    return launch_pdl(cartpole_step_kernel<T, false, true>, (unsigned)blocks, kStepThreads, st, a);
}

// Random policy (w == NULL): one env per thread for both element types (an env draws one Philox block per 128 env-steps
// for its action bits).  The fp32 env-PAIR kernel was tried for it and is slower (1.4e11 vs 2.4e11 env-steps/s at the time, 3.2e11 now): with
// ~27-step episodes a pair spends most packed steps with one slot parked for its reset.
template <typename T> int launch_rollout_random(const RolloutArgs<T> &a, cudaStream_t stream)
{
    const int64_t blocks = (a.env.n + kRolloutThreads - 1) / kRolloutThreads;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    if (a.euler) cartpole_rollout_kernel<T, true, false, true><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
    else cartpole_rollout_kernel<T, false, false, true><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
    return launch_status();
}
// Noisy variant (policy on the noisy observation): one env per thread for both element types.
template <typename T> int launch_rollout_noisy(const RolloutArgs<T> &a, cudaStream_t stream)
{
    const int64_t blocks = (a.env.n + kRolloutThreads - 1) / kRolloutThreads;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    if (a.euler) cartpole_rollout_kernel<T, true, true><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
    else cartpole_rollout_kernel<T, false, true><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
    return launch_status();
}
// fp64: one env per thread.  fp32: an env PAIR per thread on the packed FFMA2 pipe (renv_rollout_pair.cuh).
int launch_rollout(const RolloutArgs<double> &a, cudaStream_t stream)
{
    if (a.env.obs) return launch_rollout_noisy(a, stream);
    const int64_t blocks = (a.env.n + kRolloutThreads - 1) / kRolloutThreads;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    if (a.euler) cartpole_rollout_kernel<double, true><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
    else cartpole_rollout_kernel<double, false><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
    return launch_status();
}
int launch_rollout(const RolloutArgs<float> &a, cudaStream_t stream)
{
    if (a.env.obs) return launch_rollout_noisy(a, stream);
#if RENV_ROLLOUT_F32_PAIR
    const int64_t threads = (a.env.n + kPairSlots - 1) / kPairSlots;
    const int64_t blocks = (threads + kRolloutThreads - 1) / kRolloutThreads;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    if (a.euler) cartpole_rollout_pair_kernel<true><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
    else cartpole_rollout_pair_kernel<false><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
#else
    const int64_t blocks = (a.env.n + kRolloutThreads - 1) / kRolloutThreads;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    if (a.euler) cartpole_rollout_kernel<float, true><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
    else cartpole_rollout_kernel<float, false><<<(unsigned)blocks, kRolloutThreads, 0, stream>>>(a);
#endif
    return launch_status();
}

template <typename T>
int cartpole_rollout(const renv_cartpole_env *env, const renv_obs_noise *noise, const double w[4], double b, int K,
                     int integrator, int max_steps, uint64_t tick, const renv_dr_cfg *dr, double *stats,
                     unsigned long long *violations, void *stream)
{
    RolloutArgs<T> a;
    int rc = check_env<T>(env, false, &a.env);
    if (rc) return rc;
    rc = attach_noise<T>(noise, &a.env);
    if (rc) return rc;
    if (stats == nullptr) return RENV_E_NULL;
    if (w == nullptr && noise != nullptr) return RENV_E_ARG;      // the random policy does not look at observations
    if (!aligned(stats, 8)) return RENV_E_ALIGN;
    if (K <= 0 || K > (1 << 30)) return RENV_E_SIZE;      // per-thread step counters are 32-bit
    if (integrator != RENV_EULER && integrator != RENV_SEMI_IMPLICIT) return RENV_E_INTEGRATOR;
    rc = to_cfg4<T>(dr, &a.dr);
    if (rc) return rc;
    a.policy = w ? Policy<T>{ (T)w[0], (T)w[1], (T)w[2], (T)w[3], (T)b } : Policy<T>{ T(0), T(0), T(0), T(0), T(0) };
    a.K = K;
    a.euler = integrator == RENV_EULER;
    a.max_steps = max_steps;
    a.tick = tick;
    a.stats = stats;
    a.violations = violations;
    if (w == nullptr) return launch_rollout_random(a, static_cast<cudaStream_t>(stream));
    return launch_rollout(a, static_cast<cudaStream_t>(stream));
}

}  // namespace

extern "C" {

int renv_abi_version(void) { return RENV_ABI_VERSION; }

const char *renv_strerror(int code)
{
    switch (code) {
    case RENV_OK: return "ok";
    case RENV_E_NULL: return "required pointer is NULL";
    case RENV_E_ALIGN: return "pointer or ld violates the alignment contract (16 B data, 4 B byte arrays, ld % V)";
    case RENV_E_SIZE: return "invalid size (n <= 0, ld < n, K <= 0 or grid too large)";
    case RENV_E_DIM: return "dim outside [1, 32] (cart-pole needs 4)";
    case RENV_E_DRTYPE: return "Unknown dr_type";
    case RENV_E_INTEGRATOR: return "unknown integrator";
    case RENV_E_ARG: return "invalid argument";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown renv status";
    }
}

int renv_dr_sample_f32(float *out, int64_t n, const renv_dr_cfg *cfg, uint64_t seed, uint64_t sample_id0, uint32_t call,
                       unsigned long long *violations, void *stream)
{
    return dr_sample<float>(out, n, cfg, seed, sample_id0, call, violations, stream);
}
int renv_dr_sample_f64(double *out, int64_t n, const renv_dr_cfg *cfg, uint64_t seed, uint64_t sample_id0,
                       uint32_t call, unsigned long long *violations, void *stream)
{
    return dr_sample<double>(out, n, cfg, seed, sample_id0, call, violations, stream);
}

int renv_cartpole_reset_f32(const renv_cartpole_env *env, const uint8_t *mask, uint64_t tick, const renv_dr_cfg *dr,
                            unsigned long long *violations, void *stream)
{
    return cartpole_reset<float>(env, nullptr, mask, tick, dr, violations, stream);
}
int renv_cartpole_reset_f64(const renv_cartpole_env *env, const uint8_t *mask, uint64_t tick, const renv_dr_cfg *dr,
                            unsigned long long *violations, void *stream)
{
    return cartpole_reset<double>(env, nullptr, mask, tick, dr, violations, stream);
}

int renv_cartpole_step_f32(const renv_cartpole_env *env, const uint8_t *action, float *reward, uint8_t *done,
                           uint8_t *truncated, int integrator, int max_steps, int auto_reset, uint64_t tick,
                           const renv_dr_cfg *dr, unsigned long long *counters, void *stream)
{
    return cartpole_step<float, false>(env, nullptr, action, reward, done, truncated, integrator, max_steps, auto_reset,
                                       tick, dr, counters, stream);
}
int renv_cartpole_step_lean_f32(const renv_cartpole_env *env, const uint8_t *action, uint8_t *done, uint8_t *truncated,
                                int integrator, int max_steps, uint64_t tick, const renv_dr_cfg *dr,
                                unsigned long long *counters, void *stream)
{
    return cartpole_step<float, true>(env, nullptr, action, nullptr, done, truncated, integrator, max_steps, 1, tick, dr,
                                      counters, stream);
}
int renv_cartpole_step_f64(const renv_cartpole_env *env, const uint8_t *action, double *reward, uint8_t *done,
                           uint8_t *truncated, int integrator, int max_steps, int auto_reset, uint64_t tick,
                           const renv_dr_cfg *dr, unsigned long long *violations, void *stream)
{
    return cartpole_step<double, false>(env, nullptr, action, reward, done, truncated, integrator, max_steps, auto_reset, tick,
                                 dr, violations, stream);
}

int renv_cartpole_reset_noisy_f32(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *mask,
                                  uint64_t tick, const renv_dr_cfg *dr, unsigned long long *violations, void *stream)
{
    if (noise == nullptr) return RENV_E_NULL;
    return cartpole_reset<float>(env, noise, mask, tick, dr, violations, stream);
}
int renv_cartpole_reset_noisy_f64(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *mask,
                                  uint64_t tick, const renv_dr_cfg *dr, unsigned long long *violations, void *stream)
{
    if (noise == nullptr) return RENV_E_NULL;
    return cartpole_reset<double>(env, noise, mask, tick, dr, violations, stream);
}
int renv_cartpole_step_noisy_f32(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *action,
                                 float *reward, uint8_t *done, uint8_t *truncated, int integrator, int max_steps,
                                 int auto_reset, uint64_t tick, const renv_dr_cfg *dr, unsigned long long *violations,
                                 void *stream)
{
    if (noise == nullptr) return RENV_E_NULL;
    return cartpole_step<float, false>(env, noise, action, reward, done, truncated, integrator, max_steps, auto_reset, tick,
                                dr, violations, stream);
}
int renv_cartpole_step_noisy_f64(const renv_cartpole_env *env, const renv_obs_noise *noise, const uint8_t *action,
                                 double *reward, uint8_t *done, uint8_t *truncated, int integrator, int max_steps,
                                 int auto_reset, uint64_t tick, const renv_dr_cfg *dr, unsigned long long *violations,
                                 void *stream)
{
    if (noise == nullptr) return RENV_E_NULL;
    return cartpole_step<double, false>(env, noise, action, reward, done, truncated, integrator, max_steps, auto_reset, tick,
                                 dr, violations, stream);
}

int renv_cartpole_rollout_f32(const renv_cartpole_env *env, const double w[4], double b, int K, int integrator,
                              int max_steps, uint64_t tick, const renv_dr_cfg *dr, double *stats,
                              unsigned long long *violations, void *stream)
{
    return cartpole_rollout<float>(env, nullptr, w, b, K, integrator, max_steps, tick, dr, stats, violations, stream);
}
int renv_cartpole_rollout_f64(const renv_cartpole_env *env, const double w[4], double b, int K, int integrator,
                              int max_steps, uint64_t tick, const renv_dr_cfg *dr, double *stats,
                              unsigned long long *violations, void *stream)
{
    return cartpole_rollout<double>(env, nullptr, w, b, K, integrator, max_steps, tick, dr, stats, violations, stream);
}

int renv_cartpole_rollout_noisy_f32(const renv_cartpole_env *env, const renv_obs_noise *noise, const double w[4], double b,
                                    int K, int integrator, int max_steps, uint64_t tick, const renv_dr_cfg *dr,
                                    double *stats, unsigned long long *violations, void *stream)
{
    if (noise == nullptr) return RENV_E_NULL;
    return cartpole_rollout<float>(env, noise, w, b, K, integrator, max_steps, tick, dr, stats, violations, stream);
}
int renv_cartpole_rollout_noisy_f64(const renv_cartpole_env *env, const renv_obs_noise *noise, const double w[4], double b,
                                    int K, int integrator, int max_steps, uint64_t tick, const renv_dr_cfg *dr,
                                    double *stats, unsigned long long *violations, void *stream)
{
    if (noise == nullptr) return RENV_E_NULL;
    return cartpole_rollout<double>(env, noise, w, b, K, integrator, max_steps, tick, dr, stats, violations, stream);
}

int renv_cartpole_scalar_serve(renv_scalar_ctrl *ctrl, void *save, uint32_t lease_id, uint64_t lease_ns, void *stream)
{
    if (ctrl == nullptr || save == nullptr) return RENV_E_NULL;
    if (!aligned(ctrl, 64) || !aligned(save, 16)) return RENV_E_ALIGN;
    if (lease_ns == 0 || lease_ns > 1000000000ull) return RENV_E_ARG;      // a lease, not a daemon: at most 1 s idle
    cartpole_scalar_server_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(ctrl, static_cast<ScalarSave *>(save),
                                                                                 lease_id, lease_ns);
    return launch_status();
}

int renv_random_actions_u8(uint8_t *action, int64_t n, uint64_t env_id0, uint64_t seed, uint32_t step, void *stream)
{
    if (action == nullptr) return RENV_E_NULL;
    if (n <= 0) return RENV_E_SIZE;
    if (!aligned(action, 4)) return RENV_E_ALIGN;
    const int64_t threads = (n + 3) / 4;
    const int64_t blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    random_actions_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(action, n, env_id0, seed, step);
    return launch_status();
}

int renv_pack_flags_u8(const uint8_t *flags, uint32_t *bits, int64_t n, void *stream)
{
    if (flags == nullptr || bits == nullptr) return RENV_E_NULL;
    if (n <= 0) return RENV_E_SIZE;
    if (!aligned(flags, 4) || !aligned(bits, 4)) return RENV_E_ALIGN;
    const int64_t threads = (n + 31) / 32;
    const int64_t blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffLL) return RENV_E_SIZE;
    pack_flags_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(flags, bits, n);
    return launch_status();
}

}  // extern "C"
