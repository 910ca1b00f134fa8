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
#pragma once
#include "renv_philox.cuh"

namespace renv {

constexpr double kPhiMinus2 = 0.022750131948179195;      // Phi(-2)
constexpr double kPhiSpan = 0.9544997361036416;          // Phi(2) - Phi(-2)
constexpr double kGaussianFloor = 0.1;                   // random_env.py:181 (hard-coded)

enum DrType : int { kDrNone = 0, kDrUniform = 1, kDrTruncnorm = 2, kDrGaussian = 3 };

// Compact 4-dim image of renv_dr_cfg for the cart-pole kernels (passed by value as a kernel parameter).
struct DrCfg4 {
    int dr_type;
    int dim;
    double a[4], b[4], lb[4];
};

template <typename T> struct Num;
template <> struct Num<float> {
    __device__ static __forceinline__ float affine(float scale, float u, float off) { return fmaf(scale, u, off); }
    __device__ static __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static __forceinline__ float ndtri(float p) { return normcdfinvf(p); }
};
template <> struct Num<double> {
    // separately rounded multiply and add: the numpy expression lo + (hi-lo)*u, bit for bit
    __device__ static __forceinline__ double affine(double scale, double u, double off)
    {
        return __dadd_rn(off, __dmul_rn(scale, u));
    }
    __device__ static __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ static __forceinline__ double ndtri(double p) { return normcdfinv(p); }
};

// Fills out[k] for dims d = j*P + k < dim.  Returns the number of gaussian violations in this block.
template <typename T, typename Cfg>
__device__ __forceinline__ unsigned sample_dim_block(const Cfg &cfg, uint64_t seed, uint64_t id, uint64_t tick,
                                                     uint32_t purpose, int j, T *out)
{
    constexpr int P = Pack<T>::kPerBlock;
    const int d0 = j * P;
    unsigned violations = 0;
    if (cfg.dr_type == kDrUniform) {
        T u[P];
        Pack<T>::uniforms(draw_block(seed, id, tick, purpose, (uint32_t)j), u);
#pragma unroll
        for (int k = 0; k < P; ++k) {
            if (d0 + k < cfg.dim) {
                const T lo = (T)cfg.a[d0 + k], hi = (T)cfg.b[d0 + k];
                out[k] = Num<T>::affine(Num<T>::sub(hi, lo), u[k], lo);
            }
        }
    } else if (cfg.dr_type == kDrTruncnorm || cfg.dr_type == kDrGaussian) {
        const bool tn = cfg.dr_type == kDrTruncnorm;
        unsigned pending = 0;
#pragma unroll
        for (int k = 0; k < P; ++k)
            if (d0 + k < cfg.dim) pending |= 1u << k;
        for (int t = 0; t < 3 && pending; ++t) {
            const uint4 r = draw_block(seed, id, tick, purpose, (uint32_t)(t * 16 + j));
            T z[P];
            if (tn) {
                Pack<T>::uniforms(r, z);
#pragma unroll
                for (int k = 0; k < P; ++k)
                    z[k] = Num<T>::ndtri(Num<T>::affine((T)kPhiSpan, z[k], (T)kPhiMinus2));
            } else {
                Pack<T>::normals(r, z);
            }
#pragma unroll
            for (int k = 0; k < P; ++k) {
                if (pending & (1u << k)) {
                    const T x = Num<T>::affine((T)cfg.b[d0 + k], z[k], (T)cfg.a[d0 + k]);
                    const T floor_k = tn ? (T)cfg.lb[d0 + k] : (T)kGaussianFloor;
                    if (!(x < floor_k)) {        // the reference's loop condition is `obs < bound`
                        out[k] = x;
                        pending &= ~(1u << k);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < P; ++k) {
            if (pending & (1u << k)) {
                out[k] = tn ? (T)cfg.lb[d0 + k] : (T)kGaussianFloor;
                if (!tn) ++violations;
            }
        }
    }
    return violations;
}

}  // namespace renv
