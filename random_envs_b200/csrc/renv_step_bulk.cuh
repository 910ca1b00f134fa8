// Single step, fp32, auto-reset: the same arithmetic as cartpole_step_kernel<float, true>, but every global access of
// the main path is a BULK ASYNC COPY (cp.async.bulk, the 1-D form of TMA: no tensor map) between HBM and a
// shared-memory tile of 1024 envs, completed on an mbarrier:
//
//   load   4 state rows (4 x 4 KB) + xi rows (16 KB) + elapsed (4 KB) + action (1 KB)   -> 37 KB in, 7 copies
//   store  4 state rows + elapsed (4 KB) + reward (4 KB) + done (1 KB) [+ truncated]    -> 25 KB out, 7-8 copies
//
// instead of 10 LDG.E.128 + 8 STG per thread.  Each copy is one contiguous 1-16 KB burst issued by one thread, so
// the memory system sees 15 long streams per CTA rather than ~50 512-byte warp requests; the LSU/address pipes are
// idle and the threads only touch shared memory.  Finished envs are reset inside the tile (state, elapsed) before the
// store; their xi row goes straight to HBM (one 16-byte store per reset) because writing the whole xi tile back would
// add 16 B/env-step of traffic.  Results are bit-identical to the LDG/STG kernel (sha256 of 60 steps at 1024 / 5000 /
// 2^20 envs, profiles/exp/step_hash.py).
//
// MEASURED (B200, round 1) and therefore NOT the default (compile with -DRENV_STEP_F32_BULK=1 to select it):
//   2^26 envs 5954 vs 5895 GB/s (+1 %), 2^24 envs 5673 vs 5741 (-1 %), bench headline (4 x 2^20, parallel graph
//   branches) 0.929 vs 0.911 of the measured copy peak (+2 %), but a lone 2^20-env launch 14.9 vs 14.2 us and the
//   single-chain graph 6.0e10 vs 7.6e10 env-steps/s: a tile is load -> compute -> store with nothing overlapped
//   inside the CTA, so launch-to-launch latency grows.  Both kernels sit at the DRAM limit (ncu: 5.7 TB/s real
//   traffic); the access-path change moves nothing that matters, which is itself the finding.
#pragma once
#include "renv_kernels.cuh"

namespace renv {

constexpr int kBulkTile = 1024;          // envs per CTA
constexpr int kBulkThreads = 256;        // 4 envs per thread, strided by 256 (conflict-free LDS/STS)

struct __align__(128) StepTile {
    float state[4][kBulkTile];
    float xi[kBulkTile * 4];
    int32_t elapsed[kBulkTile];
    float reward[kBulkTile];
    uint8_t action[kBulkTile];
    uint8_t done[kBulkTile];
    uint8_t truncated[kBulkTile];
    uint16_t list[kBulkTile];
    unsigned long long bar;
    unsigned count;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kBulkThreads, 4) cartpole_step_bulk_kernel(const __grid_constant__ StepArgs<float> a)
{
    __shared__ StepTile t;
    const int tid = threadIdx.x;
    const int64_t block0 = (int64_t)blockIdx.x * kBulkTile;       // the launcher only covers full tiles
    const int64_t ld = a.env.ld;
    constexpr uint32_t kBytesIn = 4 * kBulkTile * 4 + kBulkTile * 16 + kBulkTile * 4 + kBulkTile;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&t.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(&t.bar)), "r"(kBytesIn) : "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_load(t.state[c], a.env.state + c * ld + block0, kBulkTile * 4, &t.bar);
        bulk_load(t.xi, a.env.xi + 4 * block0, kBulkTile * 16, &t.bar);
        bulk_load(t.elapsed, a.env.elapsed + block0, kBulkTile * 4, &t.bar);
        bulk_load(t.action, a.action + block0, kBulkTile, &t.bar);
        t.count = 0;
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) t.reward[tid + v * kBulkThreads] = 1.0f;      // :207-212: 1.0 on every step with auto-reset
    __syncthreads();                                                          // barrier initialised + count visible
    {
        uint32_t ready = 0;
        while (!ready)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ready) : "r"(smem_addr(&t.bar)) : "memory");
    }

    const bool euler = a.euler != 0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int e = tid + v * kBulkThreads;
        State<float> st = { t.state[0][e], t.state[1][e], t.state[2][e], t.state[3][e] };
        const float4 x = *reinterpret_cast<const float4 *>(t.xi + 4 * e);
        const Xi<float> p = { x.x, x.y, x.z, x.w };
        const bool terminated = dynamics(st, p, derive(p), t.action[e], euler);
        int32_t el = t.elapsed[e] + 1;                                         // TimeLimit.step
        bool done = terminated, trunc = false;
        if (a.max_steps > 0 && el >= a.max_steps) { trunc = !terminated; done = true; }
        if (done) {
            el = 0;
            t.list[atomicAdd(&t.count, 1u)] = (uint16_t)e;
        }
        t.state[0][e] = st.x; t.state[1][e] = st.x_dot; t.state[2][e] = st.theta; t.state[3][e] = st.theta_dot;
        t.elapsed[e] = el; t.done[e] = done; t.truncated[e] = trunc;
    }
    __syncthreads();                                                          // list complete, owners' tile writes done

    // resets at full lane utilisation (see cartpole_step_kernel); the new state lands in the tile, xi in HBM
    const unsigned count = t.count;
    unsigned viol = 0;
    for (unsigned j = tid; j < count; j += kBulkThreads) {
        const int e = t.list[j];
        const int64_t i = block0 + e;
        const uint64_t id = a.env.env_id0 + (uint64_t)i;
        State<float> st;
        init_state(st, a.env.seed, id, a.tick);
        t.state[0][e] = st.x; t.state[1][e] = st.x_dot; t.state[2][e] = st.theta; t.state[3][e] = st.theta_dot;
        if (a.dr.dr_type != kDrNone) {
            Xi<float> xi = { 0.0f, 0.0f, 0.0f, 0.0f };
            viol += sample_xi(xi, a.dr, a.env.seed, id, a.tick);
            store_xi(a.env.xi, i, xi);
        }
        if (a.env.episode) atomicAdd(a.env.episode + i, 1u);
    }
    if (viol && a.violations) atomicAdd(a.violations, (unsigned long long)viol);

    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");              // generic-proxy tile writes -> async proxy
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_store(a.env.state + c * ld + block0, t.state[c], kBulkTile * 4);
        bulk_store(a.env.elapsed + block0, t.elapsed, kBulkTile * 4);
        bulk_store(a.reward + block0, t.reward, kBulkTile * 4);
        bulk_store(a.done + block0, t.done, kBulkTile);
        if (a.truncated) bulk_store(a.truncated + block0, t.truncated, kBulkTile);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // the tile must outlive the copies
    }
}

}  // namespace renv
