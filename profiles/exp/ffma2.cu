#include <cstdio>
#include <cuda_runtime.h>
#define ILP 8
__global__ void k_ffma(float *out, int iters, float a, float b) {
    float acc[ILP];
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = fmaf(acc[k], a, b);
    }
    float s = 0; for (int k = 0; k < ILP; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma_imm(float *out, int iters) {
    float acc[ILP];
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = fmaf(acc[k], 1.0000001f, 1e-7f);
    }
    float s = 0; for (int k = 0; k < ILP; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3 distinct register operands, varying b per chain
__global__ void k_ffma_3reg(float *out, int iters, float a, float b) {
    float acc[ILP], bb[ILP];
    for (int k = 0; k < ILP; ++k) { acc[k] = threadIdx.x + k; bb[k] = b * (k + 1); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = fmaf(acc[k], bb[(k + 1) % ILP], bb[k]);
    }
    float s = 0; for (int k = 0; k < ILP; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__global__ void k_ffma2(float *out, int iters, float a, float b) {
    unsigned long long acc[ILP], bb[ILP];
    for (int k = 0; k < ILP; ++k) {
        float2 v = make_float2(threadIdx.x + k, threadIdx.x - k), w = make_float2(b * (k + 1), a * (k + 1));
        acc[k] = *reinterpret_cast<unsigned long long *>(&v); bb[k] = *reinterpret_cast<unsigned long long *>(&w);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = fma2(acc[k], bb[(k + 1) % ILP], bb[k]);
    }
    float s = 0; for (int k = 0; k < ILP; ++k) { float2 v = *reinterpret_cast<float2 *>(&acc[k]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sm * 8, threads = 256, iters = 20000;
    float *out; cudaMalloc(&out, blocks * threads * 4);
    double n = (double)blocks * threads * ILP * iters;
    float ms;
    ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f); });
    printf("FFMA (a,b uniform regs): %.2f TFLOP/s\n", 2 * n / ms / 1e9);
    ms = timeit([&] { k_ffma_imm<<<blocks, threads>>>(out, iters); });
    printf("FFMA imm: %.2f TFLOP/s\n", 2 * n / ms / 1e9);
    ms = timeit([&] { k_ffma_3reg<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f); });
    printf("FFMA 3 distinct regs: %.2f TFLOP/s\n", 2 * n / ms / 1e9);
    ms = timeit([&] { k_ffma2<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f); });
    printf("FFMA2 (f32x2) 3 regs: %.2f TFLOP/s  (%.3g FFMA2 warp-instr/s)\n", 4 * n / ms / 1e9, n / 32 / ms * 1e3);
    return 0;
}
