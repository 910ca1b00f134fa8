// Domain-randomisation samplers: device-side equivalent of RandomEnv.sample_task
// (reference random_envs/random_env.py:148-190), one Philox block (4 floats / 2 doubles) at a time.
//
//   uniform    :150-151  lo + (hi-lo)*U
//   truncnorm  :153-171  X = mean + std*Phi^-1(Phi(-2) + U*(Phi(2)-Phi(-2)))  -- scipy's truncnorm.rvs IS
//                        the inverse CDF of one uniform; while X < lb redraw; after the 3rd redraw X = lb.
//                        (The reference consumes a 4th draw before giving up and discards it; with a
//                        counter-based generator an unused draw has no effect, so attempts 0..2 suffice.)
//   gaussian   :173-190  X = randn*std + mean; while X < 0.1 redraw; three failures -> the reference raises;
//                        here the dim is counted in `violations` and set to the floor, the host raises.
//
// double: library normcdfinv / log / sincospi with explicitly rounded affine maps (parity path).
// float:  the truncation to [-2, 2] keeps the inverse CDF in its central region, so Phi^-1 is one MUFU.LG2 and a
//         degree-6 polynomial (no tail branches, max abs error 2.9e-7 = 1.2 ulp at |z| = 2); Box-Muller uses the
//         MUFU log / sin / cos / rsqrt approximations (abs error ~2^-21, far below anything a DR law can resolve).
#pragma once
#include "renv_philox.cuh"

namespace renv {

constexpr double kPhiMinus2 = 0.022750131948179195;      // Phi(-2)
constexpr double kPhiSpan = 0.9544997361036416;          // Phi(2) - Phi(-2)
constexpr double kGaussianFloor = 0.1;                   // random_env.py:181 (hard-coded)

enum DrType : int { kDrNone = 0, kDrUniform = 1, kDrTruncnorm = 2, kDrGaussian = 3, kDrFullGaussian = 4 };

// Compact 4-dim image of renv_dr_cfg for the cart-pole kernels (passed by value as a kernel parameter).
// fullgaussian (random_env.py:192-198): a = mean in the normalised [0,4] space, b / lb = search-bound lo / hi,
// factor = any F with F F^T = cov (row-major 4x4).
struct DrCfg4 {
    int dr_type;
    int dim;
    double a[4], b[4], lb[4];
    double factor[16];
};

template <typename T> struct Num;
template <> struct Num<float> {
    __device__ static __forceinline__ float affine(float scale, float u, float off) { return fmaf(scale, u, off); }
    __device__ static __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    // z = Phi^-1(Phi(-2) + u * (Phi(2) - Phi(-2))), u in [0, 1).  With x = 2p - 1 in [-X, X], X = Phi(2) - Phi(-2),
    // z = sqrt(2) erfinv(x) = x * g(w), w = -ln(1 - x^2) in [0, 2.42]; g is fitted (Chebyshev nodes, degree 6 in
    // t = w - 1.25).  1 - |x| is formed from s = min(u, 1-u) (exact in fp32) to avoid cancellation near |z| = 2.
    __device__ static __forceinline__ float tn_z(float u)
    {
        const float X = (float)kPhiSpan;
        const float s = fminf(u, __fsub_rn(1.0f, u));
        const float a = fmaf(2.0f * X, s, 1.0f - X);              // 1 - |x|
        const float b = __fsub_rn(2.0f, a);                        // 1 + |x|
        const float t = fmaf(__log2f(__fmul_rn(a, b)), -0.69314718056f, -1.25f);
        float g = -6.496782899e-06f;
        g = fmaf(g, t, 4.291892561e-05f);
        g = fmaf(g, t, 2.023686373e-04f);
        g = fmaf(g, t, -3.186480695e-03f);
        g = fmaf(g, t, 3.552299202e-03f);
        g = fmaf(g, t, 3.528738932e-01f);
        g = fmaf(g, t, 1.682294103e+00f);
        const float z = __fmul_rn(g, __fsub_rn(1.0f, a));
        return u >= 0.5f ? z : -z;
    }
    // 4 standard normals from one Philox block: Box-Muller on the pairs (x,y) and (z,w), both branches used
    __device__ static __forceinline__ void normals(uint4 r, float z[4])
    {
        const float two_pi = 6.283185307179586f;
        float rad = __fsqrt_rn(-2.0f * __logf(u01_open0(r.x)));
        float ang = two_pi * (u01(r.y) - 0.5f);
        z[0] = rad * __cosf(ang); z[1] = rad * __sinf(ang);
        rad = __fsqrt_rn(-2.0f * __logf(u01_open0(r.z)));
        ang = two_pi * (u01(r.w) - 0.5f);
        z[2] = rad * __cosf(ang); z[3] = rad * __sinf(ang);
    }
};
template <> struct Num<double> {
    // separately rounded multiply and add: the numpy expression lo + (hi-lo)*u, bit for bit
    __device__ static __forceinline__ double affine(double scale, double u, double off)
    {
        return __dadd_rn(off, __dmul_rn(scale, u));
    }
    __device__ static __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ static __forceinline__ double tn_z(double u) { return normcdfinv(affine(kPhiSpan, u, kPhiMinus2)); }
    __device__ static __forceinline__ void normals(uint4 r, double z[2]) { Pack<double>::normals(r, z); }
};

// fullgaussian tail (random_env.py:194-198 + denormalize_parameters :205-220): clip to [0, 4], then
// (p * (hi - lo)) / 4 + lo in that operator order.
__device__ __forceinline__ float denormalize(float x, float lo, float hi)
{
    x = fminf(fmaxf(x, 0.0f), 4.0f);
    return fmaf(__fmul_rn(x, __fsub_rn(hi, lo)), 0.25f, lo);
}
__device__ __forceinline__ double denormalize(double x, double lo, double hi)
{
    x = fmin(fmax(x, 0.0), 4.0);
    return __dadd_rn(__ddiv_rn(__dmul_rn(x, __dsub_rn(hi, lo)), 4.0), lo);
}

// Parameters of one dim block (dims j*P .. j*P+P-1) held in registers, already converted to T.
template <typename T> struct DimBlock {
    static constexpr int P = Pack<T>::kPerBlock;
    T a[P], b[P], floor[P];      // floor: lower bound (truncnorm) / 0.1 (gaussian) / unused (uniform)
    unsigned valid;              // bit k set <=> dim j*P + k exists
};

template <typename T, typename Cfg> __device__ __forceinline__ DimBlock<T> load_dim_block(const Cfg &cfg, int j)
{
    constexpr int P = Pack<T>::kPerBlock;
    DimBlock<T> blk;
    blk.valid = 0;
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int d = j * P + k;
        const bool ok = d < cfg.dim;
        blk.a[k] = ok ? (T)cfg.a[d] : T(0);
        // scale: hi - lo for uniform (rounded once, as numpy's `high - low`; fp32: times 2^-24, see uniform_affine), std otherwise
        blk.b[k] = ok ? (cfg.dr_type == kDrUniform ? Pack<T>::uniform_scale(Num<T>::sub((T)cfg.b[d], (T)cfg.a[d])) : (T)cfg.b[d]) : T(0);
        blk.floor[k] = ok ? (cfg.dr_type == kDrTruncnorm ? (T)cfg.lb[d] : (T)kGaussianFloor) : T(0);
        if (ok) blk.valid |= 1u << k;
    }
    return blk;
}

// One standard draw per dim of the block for attempt `t`: truncnorm -> TN(-2,2) by inverse CDF, gaussian -> N(0,1).
template <typename T>
__device__ __forceinline__ void standard_draws(bool tn, uint64_t seed, uint64_t id, uint64_t tick, uint32_t purpose,
                                               uint32_t slot, T *z)
{
    constexpr int P = Pack<T>::kPerBlock;
    const uint4 r = draw_block(seed, id, tick, purpose, slot);
    if (tn) {
        Pack<T>::uniforms(r, z);
#pragma unroll
        for (int k = 0; k < P; ++k) z[k] = Num<T>::tn_z(z[k]);
    } else {
        Num<T>::normals(r, z);
    }
}

// Attempt 0 for the whole block, branch-free (the common case: every dim accepted).  Returns the mask of dims
// whose draw fell below the floor and must be redrawn.  uniform: out is final, mask 0.
template <typename T>
__device__ __forceinline__ unsigned first_attempt(int dr_type, const DimBlock<T> &blk, uint64_t seed, uint64_t id,
                                                  uint64_t tick, uint32_t purpose, int j, T *out)
{
    constexpr int P = Pack<T>::kPerBlock;
    unsigned pending = 0;
    if (dr_type == kDrUniform) {
        // dims beyond `dim` have a = b = 0 and are never stored by the caller
        Pack<T>::uniform_affine(draw_block(seed, id, tick, purpose, (uint32_t)j), blk.b, blk.a, out);
    } else if (dr_type == kDrTruncnorm || dr_type == kDrGaussian) {
        T z[P];
        standard_draws<T>(dr_type == kDrTruncnorm, seed, id, tick, purpose, (uint32_t)j, z);
#pragma unroll
        for (int k = 0; k < P; ++k) {
            out[k] = Num<T>::affine(blk.b[k], z[k], blk.a[k]);
            if (out[k] < blk.floor[k]) pending |= 1u << k;      // the reference's loop condition: `obs < bound`
        }
        pending &= blk.valid;
    }
    return pending;
}

// Redraws (attempts 1 and 2) for the rejected dims only, then the reference's give-up rule.  Returns the number
// of gaussian violations.
template <typename T>
__device__ __forceinline__ unsigned redraws(int dr_type, const DimBlock<T> &blk, uint64_t seed, uint64_t id,
                                            uint64_t tick, uint32_t purpose, int j, unsigned pending, T *out)
{
    constexpr int P = Pack<T>::kPerBlock;
    const bool tn = dr_type == kDrTruncnorm;
    unsigned violations = 0;
    for (int t = 1; t < 3 && pending; ++t) {
        T z[P];
        standard_draws<T>(tn, seed, id, tick, purpose, (uint32_t)(t * 16 + j), z);
#pragma unroll
        for (int k = 0; k < P; ++k) {
            if (pending & (1u << k)) {
                const T x = Num<T>::affine(blk.b[k], z[k], blk.a[k]);
                if (!(x < blk.floor[k])) {
                    out[k] = x;
                    pending &= ~(1u << k);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < P; ++k) {
        if (pending & (1u << k)) {
            out[k] = blk.floor[k];
            if (!tn) ++violations;
        }
    }
    return violations;
}

// Fills out[k] for the valid dims of the block.  Returns the number of gaussian violations.
template <typename T>
__device__ __forceinline__ unsigned sample_dim_block(int dr_type, const DimBlock<T> &blk, uint64_t seed, uint64_t id,
                                                     uint64_t tick, uint32_t purpose, int j, T *out)
{
    const unsigned pending = first_attempt<T>(dr_type, blk, seed, id, tick, purpose, j, out);
    return pending ? redraws<T>(dr_type, blk, seed, id, tick, purpose, j, pending, out) : 0u;
}

template <typename T, typename Cfg>
__device__ __forceinline__ unsigned sample_dim_block(const Cfg &cfg, uint64_t seed, uint64_t id, uint64_t tick,
                                                     uint32_t purpose, int j, T *out)
{
    return sample_dim_block<T>(cfg.dr_type, load_dim_block<T>(cfg, j), seed, id, tick, purpose, j, out);
}

}  // namespace renv
