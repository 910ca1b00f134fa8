// Packed FP32 arithmetic of sm_100: fma.rn.f32x2 / mul.rn.f32x2 (SASS FFMA2 / FMUL2) on register pairs.
// Each lane rounds exactly like the scalar fmaf / __fmul_rn, in half the issue slots (measured on B200: FFMA2 sustains
// 64 TFLOP/s against 49 for 3-register FFMA at 0.44x the issue slots, profiles/r1/microbench.txt).
#pragma once
#include <cuda_runtime.h>

namespace renv {

using u64 = unsigned long long;

__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 splat(float v) { return pk(v, v); }

}  // namespace renv
