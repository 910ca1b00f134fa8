// Fused K-step rollout, fp32, env pairs on the packed FP32 pipe (pair_step of renv_rollout_pair.cuh) with the
// end-of-episode work served by the whole CTA instead of by the warp that owns the env.
//
// Why: with short episodes (random policy: 27 steps on average; w = [0, 0, 1, 0]: 73) the rollout is dominated by
// resets -- ~300 instructions of Philox, DR sampling and divisions -- and a warp that resets its own envs runs that
// code with 1-16 of its 64 slots active (cartpole_rollout_kernel / cartpole_rollout_pair_kernel park lanes and batch
// per warp: random policy 2.4e11 env-steps/s against 8.8e11 for a policy that survives).  A CTA of 256 threads owns
// 512 envs; a slot whose episode ends parks (NaN state, as in the pair kernel) and pushes a REQUEST into a queue in
// shared memory.  Every kPoolCheck packed steps the CTA votes (one barrier); when at least kPoolBatch requests are
// pending, threads 0 .. m-1 serve one request each -- all lanes busy whatever env the request belongs to -- and hand the
// fresh state / derived parameters / action bits back through shared memory.  Parked slots wait ~kPoolBatch / (512 p)
// steps (p = termination probability per step) instead of for the rest of their warp.
//
// The random policy's action bits (one Philox block per env and 128 steps, renv_kernels.cuh: random_action) travel the
// same way: a slot that crosses a 128-step boundary parks for a REFRESH request, so no warp ever runs Philox for one lane.
//
// Draws are keyed by (seed, env id, tick of the step that ended the episode) exactly as in the single-step kernel, so
// whichever thread serves a request, the trajectory is bit-identical to K launches of step(): tests compare with ==.
#pragma once
#include "renv_rollout_pair.cuh"

namespace renv {

#ifndef RENV_POOL_CHECK
#define RENV_POOL_CHECK 4
#endif
#ifndef RENV_POOL_BATCH
#define RENV_POOL_BATCH 128
#endif
#ifndef RENV_POOL_CTAS
#define RENV_POOL_CTAS 3
#endif
constexpr int kPoolCheck = RENV_POOL_CHECK;     // packed steps between two CTA votes
constexpr int kPoolBatch = RENV_POOL_BATCH;     // pending requests that trigger a service pass
constexpr int kPoolSlots = 2 * kRolloutThreads; // envs per CTA; slot sl = k * 256 + tid is env base + sl
enum PoolKind : uint32_t { kPoolReset = 0, kPoolRefresh = 1 };

struct PoolShared {
    float4 res[4][kPoolSlots];      // per slot: [0] state, [1] -g, F/M, pml/M, l*4/3, [2].x -(l*m_p/M), [3] action bits
    uint2 queue[kPoolSlots];        // (slot | kind << 16, env-steps done when the slot parked); a slot has one request at most
    unsigned tail;                  // push counter (atomic)
    unsigned votes[2][kRolloutThreads / 32];    // per warp: pushes << 16 | threads with an active slot; by vote parity
};

// One request, served by whichever thread drew it: RandomCartPoleEnv.reset (+ set_random_task) at the clock tick of the
// step that ended the episode, and the action bits of the step the slot resumes at.  Returns the gaussian-DR failures.
// Not inlined: the ~300 instructions and their temporaries stay out of the register budget of the stepping loop.
#ifndef RENV_POOL_SERVE_INLINE
#define RENV_POOL_SERVE_INLINE 0
#endif
#if RENV_POOL_SERVE_INLINE
#define RENV_POOL_SERVE_ATTR __forceinline__
#else
#define RENV_POOL_SERVE_ATTR __noinline__
#endif
template <bool kRandom>
__device__ RENV_POOL_SERVE_ATTR unsigned pool_serve(const RolloutArgs<float> &a, PoolShared &sh, const uint2 e, const int64_t base)
{
    const int sl = (int)(e.x & 0xffffu);
    const int steps_done = (int)e.y;
    const int64_t i = base + sl;
    const uint64_t id = a.env.env_id0 + (uint64_t)i;
    const uint32_t tick32 = (uint32_t)a.tick;
    unsigned viol = 0;
    if ((e.x >> 16) == kPoolReset) {
        const uint64_t tick = a.tick + (uint64_t)(steps_done - 1);
        if (a.dr.dr_type != kDrNone) {
            Xi<float> p = { 0.0f, 0.0f, 0.0f, 0.0f };
            viol = sample_xi(p, a.dr, a.env.seed, id, tick);
            store_xi(a.env.xi, i, p);
            PairSlot t;
            slot_params(t, p);
            sh.res[1][sl] = make_float4(t.ng, t.fot, t.pmlot, t.l43);
            sh.res[2][sl].x = t.nlpm;
        }
        State<float> st;
        init_state(st, a.env.seed, id, tick);
        sh.res[0][sl] = make_float4(st.x, st.x_dot, st.theta, st.theta_dot);
        if (a.env.episode) atomicAdd(a.env.episode + i, 1u);
    }
    if (kRandom) {
        const uint4 r = draw_block(a.env.seed, id, (uint64_t)((tick32 + (uint32_t)steps_done) >> 7), kAction, 0);
        sh.res[3][sl] = make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
    }
    return viol;
}

template <bool kEuler, bool kRandom>
__global__ void __launch_bounds__(kRolloutThreads, RENV_POOL_CTAS)
cartpole_rollout_pool_kernel(const __grid_constant__ RolloutArgs<float> a)
{
    __shared__ PoolShared sh;
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * kPoolSlots;
    const int64_t ld = a.env.ld, n = a.env.n;
    const int K = a.K;
    const int32_t limit = a.max_steps > 0 ? a.max_steps : 0x7fffffff;
    const uint32_t tick32 = (uint32_t)a.tick;
    const bool resample = a.dr.dr_type != kDrNone;
    const float nan = __int_as_float(0x7fc00000);

    unsigned long long sum_r2 = 0;
    unsigned sum_r = 0, episodes = 0, viol = 0;
    float min_r = __int_as_float(0x7f800000), max_r = __int_as_float(0xff800000);
    int c = 0;                                  // packed steps executed by this thread
    unsigned npush = 0;                         // requests this thread pushed since the last vote
    unsigned head = 0, tail = 0;                // queue window, identical in every thread (tail follows the votes)
    int parity = 0;
    PairSlot q[2];
    uint4 bits[2];                              // random policy: the slot's 128 action bits
    unsigned pos[2];                            // queue position of the slot's pending request
    bool refresh[2];                            // ... and its kind

    if (tid == 0) sh.tail = 0u;
    __syncthreads();

    auto deactivate = [&](PairSlot &s) { s.x = nan; s.xd = nan; s.th = nan; s.thd = nan; s.L = kInactive; };
    auto activate = [&](PairSlot &s, int steps_done, int elapsed) {
        s.B = steps_done - c; s.E = elapsed - c;
        long long lk = min((long long)K - s.B, (long long)limit - s.E);
        // random policy: the action bits in hand end at the next multiple of 128 of the step clock
        if (kRandom) lk = min(lk, (long long)c + 128 - (long long)((tick32 + (uint32_t)steps_done) & 127u));
        s.L = (int)min(lk, (long long)(kInactive - 1));
    };
    auto store_env = [&](const PairSlot &s, int64_t i, int elapsed) {
        a.env.state[i] = s.x; a.env.state[ld + i] = s.xd; a.env.state[2 * ld + i] = s.th; a.env.state[3 * ld + i] = s.thd;
        a.env.elapsed[i] = elapsed;
    };
    auto push = [&](int k, int sl, uint32_t kind, int steps_done) {
        pos[k] = atomicAdd(&sh.tail, 1u);
        sh.queue[pos[k] & (kPoolSlots - 1)] = make_uint2((unsigned)sl | kind << 16, (unsigned)steps_done);
        refresh[k] = kind == kPoolRefresh;
        npush += 1u;
    };
    // slot k reached `terminated || c >= L`: the episode ended (statistics, RESET request), the env has done its K steps
    // (store it), or its action bits ran out (REFRESH request; the state waits in shared memory)
    auto on_event = [&](int k, bool terminated) {
        PairSlot &s = q[k];
        const int sl = k * kRolloutThreads + tid;
        const int steps_done = s.B + c, el = s.E + c;
        if (terminated || el >= limit) {
            const float ret = (float)el;
            episodes += 1; sum_r += (unsigned)el; sum_r2 += (unsigned long long)el * (unsigned)el;   // reward is 1.0/step
            min_r = fminf(min_r, ret); max_r = fmaxf(max_r, ret);
            s.parked = steps_done; s.E = 0;
            push(k, sl, kPoolReset, steps_done);
        } else if (steps_done >= K) {
            store_env(s, base + sl, el);
        } else {
            sh.res[0][sl] = make_float4(s.x, s.xd, s.th, s.thd);
            s.parked = steps_done; s.E = el;
            push(k, sl, kPoolRefresh, steps_done);
        }
        deactivate(s);
    };
    // the owner takes a served request back
    auto pickup = [&](int k) {
        PairSlot &s = q[k];
        const int sl = k * kRolloutThreads + tid;
        const float4 st = sh.res[0][sl];
        s.x = st.x; s.xd = st.y; s.th = st.z; s.thd = st.w;
        if (!refresh[k] && resample) {
            const float4 p = sh.res[1][sl];
            s.ng = p.x; s.fot = p.y; s.pmlot = p.z; s.l43 = p.w; s.nlpm = sh.res[2][sl].x;
        }
        if (kRandom) {
            const float4 r = sh.res[3][sl];
            bits[k] = make_uint4(__float_as_uint(r.x), __float_as_uint(r.y), __float_as_uint(r.z), __float_as_uint(r.w));
        }
        const int steps_done = s.parked;
        s.parked = -1;
        if (steps_done >= K) { store_env(s, base + sl, 0); deactivate(s); }     // the episode ended on the launch's last step
        else activate(s, steps_done, s.E);
    };

    // ---- load + step 0 (scalar: an injected state may have any angle; from step 1 on |theta| <= 0.2095) -------
    const Policy<float> policy = { a.policy.w0, a.policy.w1, a.policy.w2, a.policy.w3, a.policy.b };
    bool term0[2], live[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        PairSlot &s = q[k];
        const int64_t i = base + k * kRolloutThreads + tid;
        live[k] = i < n;
        term0[k] = false;
        s.parked = -1; s.B = 0; s.E = 0;
        s.ng = -1.0f; s.fot = 1.0f; s.pmlot = 1.0f; s.l43 = 1.0f; s.nlpm = -0.5f;
        bits[k] = make_uint4(0u, 0u, 0u, 0u); pos[k] = 0u; refresh[k] = false;
        if (live[k]) {
            State<float> st = { a.env.state[i], a.env.state[ld + i], a.env.state[2 * ld + i], a.env.state[3 * ld + i] };
            const Xi<float> p = load_xi(a.env.xi, i);
            slot_params(s, p);
            int action;
            if (kRandom) {
                bits[k] = draw_block(a.env.seed, a.env.env_id0 + (uint64_t)i, (uint64_t)(tick32 >> 7), kAction, 0);
                action = (int)action_bit(bits[k], tick32);
            } else {
                action = policy_action(policy, st);
            }
            term0[k] = dynamics<false>(st, p, derive(p), action, kEuler);
            s.x = st.x; s.xd = st.x_dot; s.th = st.theta; s.thd = st.theta_dot;
            activate(s, 0, a.env.elapsed[i]);
        } else {
            deactivate(s);
        }
    }
    c = 1;
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (live[k] && (term0[k] || c >= q[k].L)) on_event(k, term0[k]);

    const PairConsts kc = { splat(policy.w0), splat(policy.w1), splat(policy.w2), splat(policy.w3), splat(policy.b),
                            splat(1.0f / 120.0f), splat(-1.0f / 6.0f), splat(-1.0f / 720.0f), splat(1.0f / 24.0f),
                            splat(-0.5f), splat(1.0f), splat((float)kTau), splat(-(float)kTau) };
    const float xthr = (float)kXThreshold, ththr = (float)kThetaThreshold;

    for (;;) {
#pragma unroll
        for (int u = 0; u < kPoolCheck; ++u) {
            if (kRandom) {
                const uint32_t now = tick32 + (uint32_t)c;
                pair_step<kEuler, true>(q[0], q[1], kc, action_bit(bits[0], now + (uint32_t)q[0].B),
                                        action_bit(bits[1], now + (uint32_t)q[1].B));
            } else {
                pair_step<kEuler>(q[0], q[1], kc);
            }
            c += 1;
            const bool t0 = fabsf(q[0].x) > xthr || fabsf(q[0].th) > ththr;
            const bool t1 = fabsf(q[1].x) > xthr || fabsf(q[1].th) > ththr;
            const bool e0 = t0 || c >= q[0].L, e1 = t1 || c >= q[1].L;
            if (e0 || e1) {
                if (e0) on_event(0, t0);
                if (e1) on_event(1, t1);
            }
        }
        // ---- CTA vote: how many requests were pushed, does anybody still step -------------------------------
        const unsigned mine = npush << 16 | ((q[0].L != kInactive || q[1].L != kInactive) ? 1u : 0u);
        const unsigned warp_sum = __reduce_add_sync(0xffffffffu, mine);
        if ((tid & 31) == 0) sh.votes[parity][tid >> 5] = warp_sum;
        __syncthreads();
        unsigned total = 0;
#pragma unroll
        for (int w = 0; w < kRolloutThreads / 32; ++w) total += sh.votes[parity][w];
        parity ^= 1; npush = 0u;
        tail += total >> 16;
        const bool stepping = (total & 0xffffu) != 0u;
        unsigned pending = tail - head;
        if (pending == 0u && !stepping) break;
        if (pending >= (unsigned)kPoolBatch || (pending != 0u && !stepping)) {
            do {
                const unsigned m = min(pending, (unsigned)kRolloutThreads);
                if ((unsigned)tid < m) viol += pool_serve<kRandom>(a, sh, sh.queue[(head + (unsigned)tid) & (kPoolSlots - 1)], base);
                __syncthreads();
                head += m;
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (q[k].parked >= 0 && (int)(pos[k] - head) < 0) pickup(k);
                pending = tail - head;
            } while (pending >= (unsigned)kPoolBatch);
        }
    }
    rollout_publish(a.stats, a.violations, episodes, sum_r, sum_r, sum_r2, min_r, max_r, viol);
}

}  // namespace renv
