#include <cstdio>
#include <cuda_runtime.h>
#include "../../random_envs_b200/csrc/renv_philox.cuh"
using namespace renv;
template <int ILP, int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, float *fout, int iters, uint64_t seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0; float facc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < ILP; ++q) {
            uint4 r = philox4x32_10(make_uint4(t, q, it, 0x03000000u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            if (MODE == 0) acc ^= r.x ^ r.y ^ r.z ^ r.w;
            else if (MODE == 1) { facc += u01(r.x) + u01(r.y) + u01(r.z) + u01(r.w); }
            else { facc += __uint_as_float((r.x >> 9) | 0x3f800000u) + __uint_as_float((r.y >> 9) | 0x3f800000u) + __uint_as_float((r.z >> 9) | 0x3f800000u) + __uint_as_float((r.w >> 9) | 0x3f800000u); }
        }
    }
    out[t] = acc; fout[t] = facc;
}
// raw op throughput
template <int OP> __global__ void __launch_bounds__(256) ops(uint32_t *out, int iters, uint32_t m) {
    uint32_t a[8];
    for (int q = 0; q < 8; ++q) a[q] = threadIdx.x * 8 + q + m;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (OP == 0) { a[q] = __umulhi(a[q], 0xD2511F53u) ^ (a[q] * 0xD2511F53u); }     // IMAD.WIDE + LOP
            if (OP == 1) { a[q] = a[q] ^ a[(q + 1) & 7] ^ m; }                               // LOP3
            if (OP == 2) { a[q] = __float_as_uint((float)(a[q] >> 8)) ; }                      // SHF + I2FP
            if (OP == 3) { a[q] = a[q] * 0xD2511F53u + m; }                                    // IMAD
            if (OP == 4) { a[q] = __umulhi(a[q], 0xD2511F53u) + m; }                          // IMAD.HI
        }
    }
    uint32_t s = 0; for (int q = 0; q < 8; ++q) s ^= a[q];
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
    int blocks = sm * 8, threads = 256, iters = 2000;
    uint32_t *out; float *fout; cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&fout, blocks * threads * 4);
    double thr = (double)blocks * threads;
#define RUN(ILP, MODE) { float ms = timeit([&] { k<ILP, MODE><<<blocks, threads>>>(out, fout, iters, 12345); }); \
    printf("philox ILP=%d mode=%d: %.3e blocks/s = %.0f GB/s of fp32 output\n", ILP, MODE, thr * ILP * iters / ms * 1e3, thr * ILP * iters / ms * 1e3 * 16 / 1e9); }
    RUN(1, 0) RUN(2, 0) RUN(4, 0) RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(2, 2)
#define OPS(OP, name, per) { float ms = timeit([&] { ops<OP><<<blocks, threads>>>(out, iters * 4, 7); }); \
    printf("%s: %.2f warp-instr/clk/SMSP-equivalent (at 1.9 GHz) [%d instr per op]\n", name, thr / 32 * 8 * iters * 4 * per / ms * 1e3 / (sm * 4 * 1.9e9), per); }
    OPS(0, "IMAD.WIDE+LOP", 2) OPS(1, "LOP3", 1) OPS(2, "SHF+I2FP", 2) OPS(3, "IMAD", 1) OPS(4, "IMAD.HI", 1)
    return 0;
}
