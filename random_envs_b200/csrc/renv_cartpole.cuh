// Cart-pole dynamics, termination, reward and reset as device functions shared by the single-step
// and the fused-rollout kernels (so both produce bit-identical trajectories).
//
// Reference: random_envs/random_cartpole.py
//   constants :74-86, dynamics :176-185, integrators :187-196, termination :200-205,
//   reward / steps_beyond_done :207-222, reset :226-229, set_task :157-166.
//
// Two arithmetic contracts, one per element type:
//   double  the PARITY path.  Every operation is an explicitly rounded __d*_rn intrinsic in the
//           reference's operator order, so nvcc can neither contract to FMA nor reassociate.  The three
//           divisions by total_mass share ONE correctly rounded reciprocal and are finished with Markstein's
//           FMA correction, which returns the correctly rounded IEEE quotient (same bits as `/`, 3 instructions
//           instead of ~12 each).  The only differences from CPython are sin/cos (a Taylor polynomial on
//           |theta| <= 0.25 that agrees with glibc's result on 99.6 % of arguments and is off by 1 ulp on the
//           rest; CUDA's <= 2-ulp sincos beyond) and pow(x,2) vs x*x.
//   float   the THROUGHPUT path.  Same formulae with hand-placed FMAs, 1/total_mass hoisted, MUFU reciprocals
//           and a small-angle sin/cos polynomial; written with intrinsics as well so the result does not
//           depend on which kernel inlines it (fused rollout == repeated single step, bit for bit).
#pragma once
#include "renv_dr.cuh"

namespace renv {

// random_cartpole.py:74-86
constexpr double kForceMag = 10.0;
constexpr double kPolemassLength = 0.05;   // pole_mass*pole_length frozen at construction; set_task never refreshes it
constexpr double kTau = 0.02;
constexpr double kXThreshold = 2.4;
constexpr double kThetaThreshold = 0.20943951023931953;   // 12 * 2 * math.pi / 360
constexpr double kFourThirds = 4.0 / 3.0;

template <typename T> struct Xi { T gravity, cart_mass, pole_mass, pole_length; };
template <typename T> struct State { T x, x_dot, theta, theta_dot; };

// Loop-invariant (per episode) quantities of the float path.
template <typename T> struct Derived;
template <> struct Derived<double> {
    double total_mass;          // m_p + m_c  (:166)
    double inv_total;           // RN(1 / total_mass), correctly rounded
};
template <> struct Derived<float> {
    float force_over_total;     // force_mag / M
    float pml_over_total;       // polemass_length / M
    float len_four_thirds;      // l * 4/3
    float len_pm_over_total;    // l * m_p / M
};

// MUFU.RCP (<= 1 ulp): the float path trades the last bit of 1/x for one instruction instead of ~8.
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// sin/cos for the float path.  With auto-reset |theta| <= 0.2095 at the start of every step (an env beyond the
// threshold was reset), so the odd/even Taylor polynomials of degree 5/6 are exact to < 0.5 ulp on |x| <= 0.25
// (first dropped terms: x^7/7! = 1.2e-8 vs sin(0.25) = 0.247, x^8/8! = 3.8e-10 vs cos ~ 0.97) in 8 FMA-pipe
// instructions; anything larger (only reachable from an injected state or when stepping past `done` without
// auto-reset) takes CUDA's full-range sincosf.
constexpr float kSmallAngle = 0.25f;
template <bool kKnownSmall = false>
__device__ __forceinline__ void sincos_small(float x, float *sn, float *cs)
{
    if (kKnownSmall || fabsf(x) <= kSmallAngle) {   // kKnownSmall: the caller guarantees it (no branch at all)
        const float x2 = __fmul_rn(x, x);
        const float ps = fmaf(x2, 1.0f / 120.0f, -1.0f / 6.0f);
        *sn = fmaf(__fmul_rn(x, x2), ps, x);
        float pc = fmaf(x2, -1.0f / 720.0f, 1.0f / 24.0f);
        pc = fmaf(x2, pc, -0.5f);
        *cs = fmaf(x2, pc, 1.0f);
    } else {
        sincosf(x, sn, cs);
    }
}

// sin/cos for the double path on |x| <= 0.25: Taylor series through x^13 / x^12 (first dropped terms
// x^15/15! < 1e-21, x^14/14! < 5e-20 of the result), Horner with FMAs.  Against glibc (the reference's math.sin /
// math.cos) over 2e7 arguments: 99.6 % identical bits, the rest 1 ulp -- closer than CUDA's library sincos --
// in 14 FP64 instructions instead of ~50 with a range reduction that is never needed here.
constexpr double kSmallAngleF64 = 0.25;
// FP64 literals cost two UMOVs each at every use (sm_100 has no 64-bit immediate operand); a __constant__ table is
// read straight from the constant bank as an instruction operand.
static __constant__ double kF64[20] = {
    1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0,      // sin: 0..5
    1.0 / 479001600.0, -1.0 / 3628800.0, 1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5,                  // cos: 6..11
    kTau, kPolemassLength, kFourThirds, kForceMag, kXThreshold, kThetaThreshold, 1.0, 0.0 };               // 12..19
template <bool kKnownSmall = false>
__device__ __forceinline__ void sincos_small(double x, double *sn, double *cs)
{
    if (kKnownSmall || fabs(x) <= kSmallAngleF64) {
        const double x2 = __dmul_rn(x, x);
        double p = kF64[0];
#pragma unroll
        for (int k = 1; k < 6; ++k) p = __fma_rn(p, x2, kF64[k]);
        *sn = __fma_rn(__dmul_rn(x, x2), p, x);
        double q = kF64[6];
#pragma unroll
        for (int k = 7; k < 12; ++k) q = __fma_rn(q, x2, kF64[k]);
        *cs = __fma_rn(x2, q, kF64[18]);
    } else {
        sincos(x, sn, cs);
    }
}

__device__ __forceinline__ Derived<double> derive(const Xi<double> &p)
{
    Derived<double> d;
    d.total_mass = __dadd_rn(p.pole_mass, p.cart_mass);                                  // :166
    d.inv_total = __drcp_rn(d.total_mass);
    return d;
}

// a / total_mass, correctly rounded: with y = RN(1/b), q = RN(a y), r = a - b q (exact in an FMA), RN(q + r y) is
// the IEEE quotient (Markstein 1990; the Itanium division sequence).  tests/test_gpu_step_parity.py checks it bit
// for bit against the oracle's `/`.
__device__ __forceinline__ double div_total(double a, const Derived<double> &d)
{
    const double q = __dmul_rn(a, d.inv_total);
    const double r = __fma_rn(-q, d.total_mass, a);
    return __fma_rn(r, d.inv_total, q);
}
__device__ __forceinline__ Derived<float> derive(const Xi<float> &p)
{
    Derived<float> d;
    const float inv_total_mass = rcp_approx(__fadd_rn(p.pole_mass, p.cart_mass));
    d.force_over_total = __fmul_rn((float)kForceMag, inv_total_mass);
    d.pml_over_total = __fmul_rn((float)kPolemassLength, inv_total_mass);
    d.len_four_thirds = __fmul_rn(p.pole_length, (float)kFourThirds);
    d.len_pm_over_total = __fmul_rn(p.pole_length, __fmul_rn(p.pole_mass, inv_total_mass));
    return d;
}

// ---- dynamics: returns terminated -----------------------------------------------------------------
template <bool kKnownSmall = false>
__device__ __forceinline__ bool dynamics(State<double> &s, const Xi<double> &p, const Derived<double> &d, int action,
                                         bool euler)
{
    const double force = action == 1 ? kF64[15] : -kF64[15];                           // :177
    double sn, cs;
    sincos_small<kKnownSmall>(s.theta, &sn, &cs);                                        // :178-179
    // temp = (force + polemass_length * theta_dot**2 * sintheta) / total_mass            :183
    const double temp = div_total(
        __dadd_rn(force, __dmul_rn(__dmul_rn(kF64[13], __dmul_rn(s.theta_dot, s.theta_dot)), sn)), d);
    // thetaacc = (g*sin - cos*temp) / (l * (4/3 - m_p*cos**2/total_mass))                :184
    const double num = __dsub_rn(__dmul_rn(p.gravity, sn), __dmul_rn(cs, temp));
    const double den = __dmul_rn(
        p.pole_length, __dsub_rn(kF64[14], div_total(__dmul_rn(p.pole_mass, __dmul_rn(cs, cs)), d)));
    const double theta_acc = __ddiv_rn(num, den);
    // xacc = temp - polemass_length * thetaacc * costheta / total_mass                   :185
    const double x_acc = __dsub_rn(temp, div_total(__dmul_rn(__dmul_rn(kF64[13], theta_acc), cs), d));
    const double tau = kF64[12];
    if (euler) {                                                                         // :187-191
        s.x = __dadd_rn(s.x, __dmul_rn(tau, s.x_dot));
        s.x_dot = __dadd_rn(s.x_dot, __dmul_rn(tau, x_acc));
        s.theta = __dadd_rn(s.theta, __dmul_rn(tau, s.theta_dot));
        s.theta_dot = __dadd_rn(s.theta_dot, __dmul_rn(tau, theta_acc));
    } else {                                                                             // :192-196
        s.x_dot = __dadd_rn(s.x_dot, __dmul_rn(tau, x_acc));
        s.x = __dadd_rn(s.x, __dmul_rn(tau, s.x_dot));
        s.theta_dot = __dadd_rn(s.theta_dot, __dmul_rn(tau, theta_acc));
        s.theta = __dadd_rn(s.theta, __dmul_rn(tau, s.theta_dot));
    }
    return s.x < -kF64[16] || s.x > kF64[16] || s.theta < -kF64[17] || s.theta > kF64[17];  // :200-205
}

template <bool kKnownSmall = false>
__device__ __forceinline__ bool dynamics(State<float> &s, const Xi<float> &p, const Derived<float> &d, int action,
                                         bool euler)
{
    float sn, cs;
    sincos_small<kKnownSmall>(s.theta, &sn, &cs);
    // temp = (F + pml thd^2 sin) / M ; thetaacc = (g sin - cos temp) / (l (4/3 - m_p cos^2 / M)) ; xacc = temp - pml thetaacc cos / M
    const float push = action == 1 ? d.force_over_total : -d.force_over_total;
    const float temp = fmaf(__fmul_rn(__fmul_rn(s.theta_dot, s.theta_dot), sn), d.pml_over_total, push);
    const float num = fmaf(p.gravity, sn, -__fmul_rn(cs, temp));
    const float den = fmaf(-d.len_pm_over_total, __fmul_rn(cs, cs), d.len_four_thirds);
    const float theta_acc = __fmul_rn(num, rcp_approx(den));
    const float x_acc = fmaf(-__fmul_rn(d.pml_over_total, cs), theta_acc, temp);
    const float tau = (float)kTau;
    if (euler) {
        s.x = fmaf(tau, s.x_dot, s.x);
        s.x_dot = fmaf(tau, x_acc, s.x_dot);
        s.theta = fmaf(tau, s.theta_dot, s.theta);
        s.theta_dot = fmaf(tau, theta_acc, s.theta_dot);
    } else {
        s.x_dot = fmaf(tau, x_acc, s.x_dot);
        s.x = fmaf(tau, s.x_dot, s.x);
        s.theta_dot = fmaf(tau, theta_acc, s.theta_dot);
        s.theta = fmaf(tau, s.theta_dot, s.theta);
    }
    return s.x < -(float)kXThreshold || s.x > (float)kXThreshold || s.theta < -(float)kThetaThreshold ||
           s.theta > (float)kThetaThreshold;
}

// ---- linear policy a = [w.s + b > 0] -----------------------------------------------------------
template <typename T> struct Policy { T w0, w1, w2, w3, b; };

__device__ __forceinline__ int policy_action(const Policy<double> &q, const State<double> &s)
{
    double acc = __dmul_rn(q.w0, s.x);                      // left-to-right, separately rounded
    acc = __dadd_rn(acc, __dmul_rn(q.w1, s.x_dot));
    acc = __dadd_rn(acc, __dmul_rn(q.w2, s.theta));
    acc = __dadd_rn(acc, __dmul_rn(q.w3, s.theta_dot));
    acc = __dadd_rn(acc, q.b);
    return acc > 0.0;
}
__device__ __forceinline__ int policy_action(const Policy<float> &q, const State<float> &s)
{
    float acc = fmaf(q.w0, s.x, q.b);
    acc = fmaf(q.w1, s.x_dot, acc);
    acc = fmaf(q.w2, s.theta, acc);
    acc = fmaf(q.w3, s.theta_dot, acc);
    return acc > 0.0f;
}

// ---- reset: s0 ~ U(-0.05, 0.05)^4 (:227) and, when DR is on, xi ~ sample_task() ---------------------
__device__ __forceinline__ void init_state(State<float> &s, uint64_t seed, uint64_t id, uint64_t tick)
{
    float u[4];
    Pack<float>::uniforms(draw_block(seed, id, tick, kInit, 0), u);
    s.x = fmaf(0.1f, u[0], -0.05f);
    s.x_dot = fmaf(0.1f, u[1], -0.05f);
    s.theta = fmaf(0.1f, u[2], -0.05f);
    s.theta_dot = fmaf(0.1f, u[3], -0.05f);
}
__device__ __forceinline__ void init_state(State<double> &s, uint64_t seed, uint64_t id, uint64_t tick)
{
    double u[2], v[2];
    Pack<double>::uniforms(draw_block(seed, id, tick, kInit, 0), u);
    Pack<double>::uniforms(draw_block(seed, id, tick, kInit, 1), v);
    s.x = __dadd_rn(-0.05, __dmul_rn(0.1, u[0]));           // numpy: low + (high - low) * u
    s.x_dot = __dadd_rn(-0.05, __dmul_rn(0.1, u[1]));
    s.theta = __dadd_rn(-0.05, __dmul_rn(0.1, v[0]));
    s.theta_dot = __dadd_rn(-0.05, __dmul_rn(0.1, v[1]));
}

// ---- observation noise of the suite's "Noisy" variants: obs = state + sqrt(noise_level) * N(0, I) ----------
// (jinja/random_hopper.py:107-108, random_walker2d.py:139-140, random_humanoid.py:193-204; drawn on every
// _get_obs, i.e. after a step AND after a reset.)  Four normals keyed (seed, id, tick, kObs); `which` = 0 for the
// observation of a step, 1 for the observation of the reset that may follow it at the same tick.
__device__ __forceinline__ void add_obs_noise(const State<float> &s, float std, uint64_t seed, uint64_t id, uint64_t tick,
                                              uint32_t which, float o[4])
{
    float z[4];
    Num<float>::normals(draw_block(seed, id, tick, kObs, which * 16u), z);
    o[0] = fmaf(std, z[0], s.x); o[1] = fmaf(std, z[1], s.x_dot);
    o[2] = fmaf(std, z[2], s.theta); o[3] = fmaf(std, z[3], s.theta_dot);
}
__device__ __forceinline__ void add_obs_noise(const State<double> &s, double std, uint64_t seed, uint64_t id, uint64_t tick,
                                              uint32_t which, double o[4])
{
    double z[4];
    Pack<double>::normals(draw_block(seed, id, tick, kObs, which * 16u), z);
    Pack<double>::normals(draw_block(seed, id, tick, kObs, which * 16u + 1u), z + 2);
    o[0] = __dadd_rn(s.x, __dmul_rn(std, z[0])); o[1] = __dadd_rn(s.x_dot, __dmul_rn(std, z[1]));      // obs += std * randn
    o[2] = __dadd_rn(s.theta, __dmul_rn(std, z[2])); o[3] = __dadd_rn(s.theta_dot, __dmul_rn(std, z[3]));
}

// fullgaussian for the 4-dim cart-pole xi: x = mean + F z in the normalised space, clip, denormalise.
// Deliberately NOT inlined: it is the rarest branch of the (already cold) reset path and inlining it costs the
// fused rollout kernel 40 registers.  `cfg` points into the kernel's __grid_constant__ parameter block.
template <typename T>
__device__ __noinline__ void fullgaussian_xi(const DrCfg4<T> *cfg, uint64_t seed, uint64_t id, uint64_t tick, T *v)
{
    constexpr int P = Pack<T>::kPerBlock;
    T z[4];
#pragma unroll
    for (int j = 0; j < 4 / P; ++j) Num<T>::normals(draw_block(seed, id, tick, kXi, (uint32_t)j), z + j * P);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        T x = cfg->a[d];
#pragma unroll
        for (int k = 0; k < 4; ++k) x = Num<T>::affine(cfg->factor[d * 4 + k], z[k], x);
        v[d] = denormalize(x, cfg->b[d], cfg->floor[d]);
    }
}

template <typename T>
__device__ __forceinline__ unsigned sample_xi(Xi<T> &p, const DrCfg4<T> &cfg, uint64_t seed, uint64_t id, uint64_t tick)
{
    constexpr int P = Pack<T>::kPerBlock;
    T v[4] = { p.gravity, p.cart_mass, p.pole_mass, p.pole_length };
    unsigned violations = 0;
    if (cfg.dr_type == kDrFullGaussian) {
        fullgaussian_xi<T>(&cfg, seed, id, tick, v);
    } else {
#pragma unroll
        for (int j = 0; j < 4 / P; ++j) violations += sample_dim_block<T>(cfg, seed, id, tick, kXi, j, v + j * P);
    }
    p.gravity = v[0]; p.cart_mass = v[1]; p.pole_mass = v[2]; p.pole_length = v[3];
    return violations;
}

}  // namespace renv
