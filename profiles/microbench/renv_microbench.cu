// Micro-benchmarks behind bench.py's compute rooflines -- NOT part of the product library (librenv_b200.so exports only
// the path's entry points, include/renv.h).  Built by profiles/microbench/build.py into librenv_microbench.so.
//
//   renv_fma_peak_f32 / _f64   dependent-FMA chains, 8 independent accumulators per thread:
//                              FLOPs = 2 * blocks * threads * 8 * iters   (denominator of the fused rollout's fraction)
//   renv_philox_peak           Philox4x32-10 blocks per second with nothing else in the loop, one 16-byte store per
//                              thread at the end (the issue-rate roof of the DR samplers: 20 IMAD.WIDE per 16 output bytes)
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../random_envs_b200/csrc/renv_philox.cuh"

namespace {
constexpr int kIlp = 8;
template <typename T> __global__ void fma_peak_kernel(T *out, int iters)
{
    T acc[kIlp];
    const T a = (T)1.0000001, b = (T)1e-7;
#pragma unroll
    for (int k = 0; k < kIlp; ++k) acc[k] = (T)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kIlp; ++k) acc[k] = acc[k] * a + b;
    }
    T sum = 0;
#pragma unroll
    for (int k = 0; k < kIlp; ++k) sum += acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

__global__ void philox_peak_kernel(uint4 *out, int iters, uint64_t seed)
{
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
        const uint4 r = renv::draw_block(seed, id, (uint64_t)it, renv::kTasks, 0);
        acc.x ^= r.x; acc.y ^= r.y; acc.z ^= r.z; acc.w ^= r.w;
    }
    out[id] = acc;
}

template <typename T> int fma_peak(T *out, int blocks, int threads, int iters, void *stream)
{
    if (out == nullptr || blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0) return -1;
    fma_peak_kernel<T><<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(out, iters);
    return (int)cudaGetLastError();
}
}  // namespace

extern "C" {
int renv_fma_peak_f32(float *out, int blocks, int threads, int iters, void *stream) { return fma_peak<float>(out, blocks, threads, iters, stream); }
int renv_fma_peak_f64(double *out, int blocks, int threads, int iters, void *stream) { return fma_peak<double>(out, blocks, threads, iters, stream); }
int renv_philox_peak(void *out, int blocks, int threads, int iters, uint64_t seed, void *stream)
{
    if (out == nullptr || blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0) return -1;
    philox_peak_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint4 *>(out), iters, seed);
    return (int)cudaGetLastError();
}
}
