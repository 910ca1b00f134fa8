// Single step, fp32, auto-reset: ONE WAVE of persistent CTAs (one per SM) that stream 1-D bulk-async tiles
// (cp.async.bulk, the tensor-map-free form of TMA; SASS UBLKCP) through a shared-memory ring.
//
// Why (round-1 verdict, "one-batch throughput"): the LDG/STG kernel covers 2^20 envs with 1024 CTAs on 592 resident
// slots = 1.73 waves; a lone launch spends ~2 of its ~12 us in the turnover between waves and in the load->compute->store
// serialisation inside each CTA.  Here every SM owns ONE CTA for the whole launch and a ring of kRingStages tiles:
//
//   producer (thread 0)   arms full[s] with the tile's byte count and issues 7 bulk loads HBM -> smem
//                         (4 state rows, the xi rows, elapsed, action): at kernel entry the whole ring is in flight
//                         (148 SMs x ~190 KB = 28 MB of the 39 MB a 2^20-env step reads), and it stays kRingStages-1
//                         tiles ahead for the rest of the launch -- independent of registers or occupancy;
//   consumers (all)       wait on full[s], step their envs IN the tile (same arithmetic as cartpole_step_kernel:
//                         dynamics<float>, TimeLimit, auto-reset with CTA-level reset compaction), fence to the async
//                         proxy;
//   producer              issues the bulk stores smem -> HBM (state rows, elapsed, reward, done[, truncated]) as one
//                         bulk group, waits until the PREVIOUS tile's group has finished reading its stage and refills
//                         that stage with the tile kRingStages-1 ahead.
//
// The reward tile is a constant 1.0 written once per CTA (auto-reset: random_cartpole.py:207-212); a finished env's xi row
// goes straight to HBM (one 16-byte store per reset: writing xi tiles back would add 16 B/env-step).  Only whole tiles
// are handled here; the launcher gives the remainder (n % kRingTile envs) to the LDG/STG kernel.
// Results are bit-identical to cartpole_step_kernel<float, true> (tests/test_gpu_step_ring.py compares with ==).
#pragma once
#include "renv_kernels.cuh"

namespace renv {

#ifndef RENV_RING_TILE
#define RENV_RING_TILE 512
#endif
#ifndef RENV_RING_STAGES
#define RENV_RING_STAGES 10
#endif
#ifndef RENV_RING_THREADS
#define RENV_RING_THREADS 512
#endif
#ifndef RENV_RING_GROUPS
#define RENV_RING_GROUPS 4          // independent consumer groups: group g steps tiles g, g + G, ... on its own barrier
#endif
#ifndef RENV_RING_CTAS_PER_SM
#define RENV_RING_CTAS_PER_SM 1
#endif
#ifndef RENV_RING_EARLY_TRIGGER
#define RENV_RING_EARLY_TRIGGER 1
#endif
#ifndef RENV_RING_STORE_DEPTH
#define RENV_RING_STORE_DEPTH 4     // tiles whose bulk stores may still be reading their stage (the rest prefetch)
#endif
constexpr int kRingTile = RENV_RING_TILE;            // envs per tile
constexpr int kRingStages = RENV_RING_STAGES;
constexpr int kRingThreads = RENV_RING_THREADS;
constexpr int kRingGroups = RENV_RING_GROUPS;
constexpr int kGroupThreads = kRingThreads / kRingGroups;
constexpr int kRingPerThread = kRingTile / kGroupThreads;
static_assert(kRingThreads % kRingGroups == 0 && kGroupThreads % 32 == 0 && kRingTile % kGroupThreads == 0 &&
              kRingGroups <= 15, "consumer groups");
constexpr int kRingStoreDepth = RENV_RING_STORE_DEPTH;
static_assert(kRingStoreDepth >= 1 && kRingStoreDepth < kRingStages, "store depth");
static_assert(kRingTile % 16 == 0 && kRingTile <= 65536, "ring tile geometry");

template <typename E> struct __align__(128) RingStage {
    float state[4][kRingTile];
    float xi[kRingTile * 4];
    E elapsed[kRingTile];
    uint8_t action[kRingTile];
    uint8_t done[kRingTile];
    uint8_t truncated[kRingTile];
};
template <typename E> struct __align__(128) RingSmem {
    RingStage<E> stage[kRingStages];
    float reward[kRingTile];
    uint16_t list[kRingGroups][2][kRingTile];
    unsigned long long full[kRingStages];      // producer -> consumers: the tile's bulk loads have landed
    unsigned long long ready[kRingStages];     // consumers -> producer: the tile is stepped and may be stored
    unsigned count[kRingGroups][4];            // [g][0..2]: finished-env counters of the group's last three tiles
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
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t ready = 0;
    while (!ready)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ready) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}

// E = int32_t: the 62-byte step (elapsed int32, reward written).  E = uint16_t: the lean 54-byte step
// (EnvPtrs::elapsed16, reward == nullptr); see renv_cartpole_step_lean_f32 in include/renv.h.
//
// Threads 0 .. kRingThreads-1 are consumers; one extra warp is the producer (its lane 0 issues every bulk copy), so a
// consumer never waits for a store to drain: it arrives on ready[s] and goes on to the next tile.  The consumers form
// kRingGroups independent groups (own named barrier, own reset list): a tile's load -> step -> barrier -> resets ->
// fence chain is ~1 us of LATENCY whatever the thread count (measured: 256 or 512 threads on one tile at a time both
// give ~1 us per 512-env tile = half the rate HBM delivers), so several tiles are stepped concurrently.
template <typename E>
__global__ void __launch_bounds__(kRingThreads + 32, RENV_RING_CTAS_PER_SM)
cartpole_step_ring_kernel(const __grid_constant__ StepArgs<float> a, const int num_tiles)
{
    extern __shared__ __align__(128) unsigned char ring_raw[];
    RingSmem<E> &sm = *reinterpret_cast<RingSmem<E> *>(ring_raw);
    const int tid = threadIdx.x;
    const int64_t ld = a.env.ld;
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_tiles = (num_tiles - first + stride - 1) / stride;      // >= 1: the launcher keeps grid <= num_tiles
    E *const g_elapsed = sizeof(E) == 2 ? reinterpret_cast<E *>(a.env.elapsed16) : reinterpret_cast<E *>(a.env.elapsed);
    constexpr uint32_t kBytesIn = 4 * kRingTile * 4 + kRingTile * 16 + kRingTile * sizeof(E) + kRingTile;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRingStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&sm.full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(&sm.ready[s])), "n"(kGroupThreads));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kRingGroups * 4) sm.count[tid >> 2][tid & 3] = 0;
    if (a.reward)
        for (int e = tid; e < kRingTile; e += kRingThreads + 32) sm.reward[e] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the reward tile is read by bulk stores
    __syncthreads();                        // barriers initialised, counters and the reward tile visible
    // everything above is private to the CTA; the first global access of every thread comes after this wait
    // (programmatic dependent launch, see cartpole_step_kernel)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#if RENV_RING_EARLY_TRIGGER
    // a dependent grid cannot become resident before this CTA exits anyway (the ring takes the SM's shared memory),
    // so it may be made launchable at once: it is scheduled the moment SMs drain
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif

    if (tid >= kRingThreads) {
        // ---------------------------------------------------------------- producer: every bulk copy of the CTA
        if (tid != kRingThreads) return;
        auto issue_load = [&](int k) {
            RingStage<E> &st = sm.stage[k % kRingStages];
            unsigned long long *bar = &sm.full[k % kRingStages];
            const int64_t env0 = (int64_t)(first + k * stride) * kRingTile;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(kBytesIn) : "memory");
#pragma unroll
            for (int c = 0; c < 4; ++c) bulk_load(st.state[c], a.env.state + c * ld + env0, kRingTile * 4, bar);
            bulk_load(st.xi, a.env.xi + 4 * env0, kRingTile * 16, bar);
            bulk_load(st.elapsed, g_elapsed + env0, kRingTile * sizeof(E), bar);
            bulk_load(st.action, a.action + env0, kRingTile, bar);
        };
        const int pre = my_tiles < kRingStages ? my_tiles : kRingStages;
        for (int k = 0; k < pre; ++k) issue_load(k);
        for (int k = 0; k < my_tiles; ++k) {
            RingStage<E> &st = sm.stage[k % kRingStages];
            const int64_t env0 = (int64_t)(first + k * stride) * kRingTile;
            mbar_wait(&sm.ready[k % kRingStages], (uint32_t)(k / kRingStages) & 1u);
#pragma unroll
            for (int c = 0; c < 4; ++c) bulk_store(a.env.state + c * ld + env0, st.state[c], kRingTile * 4);
            bulk_store(g_elapsed + env0, st.elapsed, kRingTile * sizeof(E));
            if (a.reward) bulk_store(a.reward + env0, sm.reward, kRingTile * 4);
            bulk_store(a.done + env0, st.done, kRingTile);
            if (a.truncated) bulk_store(a.truncated + env0, st.truncated, kRingTile);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the stage of tile k - D is free once its store group (D groups ago) has read it out; up to D tiles of
            // stores drain while kRingStages - D tiles of loads are in flight
            const int j = k - kRingStoreDepth;
            if (j >= 0 && j + kRingStages < my_tiles) {
                asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(kRingStoreDepth) : "memory");
                issue_load(j + kRingStages);
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");           // the ring must outlive the copies
        return;
    }

    // -------------------------------------------------------------------- consumers
    const bool euler = a.euler != 0;
    const int max_steps = a.max_steps;
    const int g = tid / kGroupThreads, gtid = tid % kGroupThreads;
    unsigned bad_actions = 0, viol = 0;
    for (int k = g, it = 0; k < my_tiles; k += kRingGroups, ++it) {
        RingStage<E> &st = sm.stage[k % kRingStages];
        unsigned *const count = &sm.count[g][it % 3];
        uint16_t *const list = sm.list[g][it & 1];
        const int64_t env0 = (int64_t)(first + k * stride) * kRingTile;
        // every thread of the group has passed barrier (1) of the group's previous tile, i.e. finished the one before:
        // that tile's counter and list are free
        if (gtid == 0) sm.count[g][(it + 1) % 3] = 0;
        mbar_wait(&sm.full[k % kRingStages], (uint32_t)(k / kRingStages) & 1u);

#pragma unroll
        for (int v = 0; v < kRingPerThread; ++v) {
            const int e = gtid + v * kGroupThreads;             // strided ownership: conflict-free LDS/STS
            State<float> s = { st.state[0][e], st.state[1][e], st.state[2][e], st.state[3][e] };
            const float4 x = *reinterpret_cast<const float4 *>(st.xi + 4 * e);
            const Xi<float> p = { x.x, x.y, x.z, x.w };
            const unsigned act = st.action[e];
            bad_actions |= act > 1u;
            const bool terminated = dynamics(s, p, derive(p), (int)act, euler);
            int el = (int)st.elapsed[e] + 1;                                   // TimeLimit.step
            bool done = terminated, trunc = false;
            if (max_steps > 0 && el >= max_steps) { trunc = !terminated; done = true; }
            if (done) {
                el = 0;
                list[atomicAdd(count, 1u)] = (uint16_t)e;
            }
            st.state[0][e] = s.x; st.state[1][e] = s.x_dot; st.state[2][e] = s.theta; st.state[3][e] = s.theta_dot;
            st.elapsed[e] = (E)el; st.done[e] = done; st.truncated[e] = trunc;
        }
        asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "n"(kGroupThreads) : "memory");   // (1) list complete, tile writes done

        // resets at full lane utilisation (see cartpole_step_kernel); the new state lands in the tile, xi in HBM
        const unsigned cnt = *count;
        for (unsigned j = gtid; j < cnt; j += kGroupThreads) {
            const int e = list[j];
            const int64_t i = env0 + e;
            const uint64_t id = a.env.env_id0 + (uint64_t)i;
            State<float> s;
            init_state(s, a.env.seed, id, a.tick);
            st.state[0][e] = s.x; st.state[1][e] = s.x_dot; st.state[2][e] = s.theta; st.state[3][e] = s.theta_dot;
            if (a.dr.dr_type != kDrNone) {
                Xi<float> xi = { 0.0f, 0.0f, 0.0f, 0.0f };
                viol += sample_xi(xi, a.dr, a.env.seed, id, a.tick);
                store_xi(a.env.xi, i, xi);
            }
            if (a.env.episode) atomicAdd(a.env.episode + i, 1u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy tile writes -> async proxy
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(&sm.ready[k % kRingStages])) : "memory");
    }
#if !RENV_RING_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    if (viol && a.counters) atomicAdd(a.counters + kCounterGaussian, (unsigned long long)viol);
    if (bad_actions && a.counters) atomicAdd(a.counters + kCounterBadAction, 1ull);
}

}  // namespace renv
