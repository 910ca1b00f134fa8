// The scalar drop-in env (gym.make('RandomCartPole-v0'): one env, one Python call per step) without a kernel launch
// per step: a one-warp RESIDENT kernel serves step / reset requests that the host rings in through a word of pinned,
// device-mapped host memory and answers through the same block.
//
// Why: a launch plus a stream synchronise costs ~25 us per scalar step, 4-7x the reference's CPython step
// (random_cartpole.py:172-224, 3.7-7 us); a request that is one 4-byte PCIe write and an answer that is one 64-byte
// posted write cost 3-4 us.  The env's state, xi and steps_beyond_done live in the kernel's registers between requests.
//
// Protocol (struct renv_scalar_ctrl, include/renv.h; all in HOST memory, the kernel polls it over PCIe):
//   host   request = (seq << 8) | op             seq grows by 1 per request (24 bits); ops 0 / 1 are step(action),
//                                                 the others read their arguments from ctrl->arg (written BEFORE request)
//   device reads request until its seq is last + 1; executes; writes the results; __threadfence_system(); ack = seq
//   host   spins on ack == seq
// Look-ahead (configure request, arg_u64[5] != 0): while it waits for the next request the kernel also publishes BOTH
// possible next steps (ctrl->next[0..1], then next_seq = seq of the request they follow).  The host's next step(a)
// returns next[a] at once and rings the step in as op 8 + a without waiting (not acknowledged; one request outstanding
// at most: the host waits for next_seq == that seq before it writes another request word).  The kernel keeps the four
// two-step futures as well, so on such a request it publishes the following pair BEFORE recomputing anything: its
// turnaround is one PCIe poll plus the posted stores, and it overlaps the caller's own work.  Measured: the
// test_random_policy.py loop 9.5 -> 6.1-6.9 us per step, env.step back to back 6.0 -> 4.5 us.
// The kernel is a LEASE, not a daemon: after `lease_ns` without a request it saves its registers to `save` (device
// memory) and exits, setting ctrl->exited = lease id; the host relaunches it with the next request (the request word
// stays pending in host memory and is served by the new instance).  So a device-wide synchronise elsewhere in the
// process stalls for at most one lease, and nothing can spin forever on either side.
#pragma once
#include <cstddef>
#include "../../include/renv.h"
#include "renv_kernels.cuh"

namespace renv {

enum ScalarOp : uint32_t { kOpStep0 = 0, kOpStep1 = 1, kOpReset = 2, kOpSetState = 3, kOpSetXi = 4, kOpConfig = 5,
                           kOpExit = 6, kOpAheadStep0 = 8, kOpAheadStep1 = 9 };

// What survives between kernel instances (device memory, 512 bytes, zero-initialised by the caller once).
struct ScalarSave {
    double state[4], xi[4], noise_std;
    int32_t beyond, euler, noisy, lookahead;
    uint32_t last_seq, valid;
    uint64_t seed, tick_next;
    DrCfg4<double> dr;
};

__device__ __forceinline__ uint32_t ld_sys_u32(const volatile uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t ld_sys_u64(const void *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_sys_f64(const double *p) { return __longlong_as_double((long long)ld_sys_u64(p)); }
__device__ __forceinline__ uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

static_assert(sizeof(ScalarSave) <= RENV_SCALAR_SAVE_BYTES, "ScalarSave outgrew RENV_SCALAR_SAVE_BYTES");
static_assert(offsetof(renv_scalar_ctrl, next) == 448 && offsetof(renv_scalar_ctrl, next_seq) == 704 &&
              sizeof(renv_scalar_ctrl) == 768, "renv_scalar_ctrl layout (random_envs_b200/random_cartpole.py mirrors it)");

// std * N(0, I) of the observation made at clock `tick` after a step: add_obs_noise<double> is state + this, with the
// same two roundings (renv_cartpole.cuh), so an observation assembled from a pre-drawn vector is bit-identical.
__device__ __forceinline__ void scalar_obs_noise(double std, uint64_t seed, uint64_t tick, double nz[4])
{
    const State<double> zero = { 0.0, 0.0, 0.0, 0.0 };
    add_obs_noise(zero, std, seed, 0ull, tick, 0u, nz);       // 0 + std * z == std * z exactly
}
__device__ __forceinline__ void scalar_put_state(double *dst, const State<double> &s)
{
    *reinterpret_cast<double2 *>(dst) = make_double2(s.x, s.x_dot);
    *reinterpret_cast<double2 *>(dst + 2) = make_double2(s.theta, s.theta_dot);
}
__device__ __forceinline__ void scalar_put_obs(double *dst, const State<double> &s, const double nz[4])
{
    *reinterpret_cast<double2 *>(dst) = make_double2(__dadd_rn(s.x, nz[0]), __dadd_rn(s.x_dot, nz[1]));
    *reinterpret_cast<double2 *>(dst + 2) = make_double2(__dadd_rn(s.theta, nz[2]), __dadd_rn(s.theta_dot, nz[3]));
}
// reward, done and steps_beyond_done of a step that ends in `terminated`, as one 16-byte store (:207-222)
__device__ __forceinline__ int32_t scalar_put_tail(double *reward_field, bool terminated, int32_t beyond)
{
    double reward = 1.0;
    if (terminated) { reward = beyond < 0 ? 1.0 : 0.0; beyond = beyond < 0 ? 0 : beyond + 1; }
    int4 tail;
    tail.x = (int)(__double_as_longlong(reward) & 0xffffffffll); tail.y = (int)(__double_as_longlong(reward) >> 32);
    tail.z = terminated ? 1 : 0; tail.w = beyond;
    *reinterpret_cast<int4 *>(reward_field) = tail;
    return beyond;
}

__global__ void __launch_bounds__(32, 1)
cartpole_scalar_server_kernel(renv_scalar_ctrl *ctrl, ScalarSave *save, uint32_t lease_id, uint64_t lease_ns)
{
    if (threadIdx.x != 0) return;
    State<double> s = { save->state[0], save->state[1], save->state[2], save->state[3] };
    Xi<double> p = { save->xi[0], save->xi[1], save->xi[2], save->xi[3] };
    if (!save->valid) p = Xi<double>{ 9.8, 1.0, 0.1, 0.5 };                 // random_cartpole.py:74-78
    double noise_std = save->noise_std;
    int32_t beyond = save->valid ? save->beyond : -1;
    bool euler = save->valid ? save->euler != 0 : true;
    bool noisy = save->valid ? save->noisy != 0 : false;
    bool lookahead = save->valid ? save->lookahead != 0 : false;
    uint64_t seed = save->seed;
    uint64_t tick_next = save->valid ? save->tick_next : 0ull;     // the step clock value of the NEXT step / reset
    uint32_t last = save->valid ? save->last_seq : ld_sys_u32(&ctrl->ack);
    DrCfg4<double> dr = save->dr;
    if (!save->valid) dr.dr_type = kDrNone;

    // While the host is busy between two calls, the possible futures are computed ahead: `one[a]` = the state after
    // step(a) from s, and in look-ahead mode `two[a][b]` = the state after step(a), step(b).  When a step request
    // arrives the answer is a select -- the ~1 us of dependent fp64 arithmetic of one thread is off the round trip --
    // and in look-ahead mode the NEXT answers (two[a][*]) are published at once, before anything is recomputed.
    State<double> one[2], two[2][2];
    bool one_term[2], two_term[2][2];
    double nz[4], nz_after[4];                   // std * N(0, I) of the observations at tick_next / tick_next + 1
    bool one_valid = false, two_valid = false, published = false, nz_valid = false, nz_after_valid = false;

    uint64_t idle_since = global_timer_ns();
    for (;;) {
        if (!one_valid) {
            const Derived<double> d = derive(p);
            one[0] = s; one[1] = s;
            one_term[0] = dynamics(one[0], p, d, 0, euler);
            one_term[1] = dynamics(one[1], p, d, 1, euler);
            one_valid = true; two_valid = false; published = false;
        }
        if (lookahead && !published) {
            // renv_scalar_ctrl.next: what step(0) / step(1) return from here; next_seq names the request they follow
            if (noisy && !nz_valid) { scalar_obs_noise(noise_std, seed, tick_next, nz); nz_valid = true; }
            for (int a = 0; a < 2; ++a) {
                renv_scalar_ctrl::renv_scalar_outcome *o = &ctrl->next[a];
                scalar_put_state(o->state, one[a]);
                if (noisy) scalar_put_obs(o->obs, one[a], nz);
                scalar_put_tail(&o->reward, one_term[a], beyond);
            }
            asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(&ctrl->next_seq), "r"(last) : "memory");
            published = true;
        }
        if (lookahead && !two_valid) {
            const Derived<double> d = derive(p);
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    two[a][b] = one[a];
                    two_term[a][b] = dynamics(two[a][b], p, d, b, euler);
                }
            two_valid = true;
        }
        if (lookahead && noisy && !nz_after_valid) {
            scalar_obs_noise(noise_std, seed, tick_next + 1, nz_after);
            nz_after_valid = true;
        }

        uint32_t r, seq, op;
        bool expired = false;
        for (;;) {
            r = ld_sys_u32(&ctrl->request);
            seq = r >> 8; op = r & 0xffu;
            if (seq == ((last + 1u) & 0xffffffu)) break;
            if (global_timer_ns() - idle_since > lease_ns) { expired = true; break; }
        }
        if (expired) break;
        unsigned viol = 0;
        const bool is_step = op <= kOpStep1 || op == kOpAheadStep0 || op == kOpAheadStep1;
        const bool ahead = op >= kOpAheadStep0;
        if (!is_step) __threadfence_system();           // the arguments were written before the request word
        if (is_step) {
            // RandomCartPoleEnv.step, no auto-reset, TimeLimit left to the gym wrapper (as in the reference stack)
            const int a = (int)(op & 1u);
            s = one[a];
            const bool terminated = one_term[a];
            if (!ahead) {
                if (noisy) {                    // Noisy variant: obs = state + std N(0, I), keyed by this step's clock value
                    if (!nz_valid) { scalar_obs_noise(noise_std, seed, tick_next, nz); nz_valid = true; }
                    scalar_put_obs(ctrl->obs, s, nz);
                }
                beyond = scalar_put_tail(&ctrl->reward, terminated, beyond);
            } else if (terminated) {
                beyond = beyond < 0 ? 0 : beyond + 1;
            }
            tick_next += 1;
            // the futures move one level up
            if (two_valid) {
                one[0] = two[a][0]; one[1] = two[a][1]; one_term[0] = two_term[a][0]; one_term[1] = two_term[a][1];
                one_valid = true;
            } else {
                one_valid = false;
            }
            two_valid = false; published = false;
            if (nz_after_valid) { nz[0] = nz_after[0]; nz[1] = nz_after[1]; nz[2] = nz_after[2]; nz[3] = nz_after[3]; }
            nz_valid = nz_after_valid; nz_after_valid = false;
        } else if (op == kOpReset) {
            // arg_u64: [0] tick, [1] resample xi (set_random_task), [2] sample_task call index, [3] dr seed
            const uint64_t tick = ld_sys_u64(&ctrl->arg_u64[0]), resample = ld_sys_u64(&ctrl->arg_u64[1]);
            if (resample && dr.dr_type != kDrNone && dr.dr_type != kDrFullGaussian) {
                // RandomEnv.sample_task (random_env.py:148-190): the draws of renv_dr_sample_f64 for sample 0 of this call
                const uint64_t dr_call = ld_sys_u64(&ctrl->arg_u64[2]), dr_seed = ld_sys_u64(&ctrl->arg_u64[3]);
                double v[4] = { p.gravity, p.cart_mass, p.pole_mass, p.pole_length };
                viol += sample_dim_block<double>(dr, dr_seed, 0ull, dr_call, kTasks, 0, v);
                viol += sample_dim_block<double>(dr, dr_seed, 0ull, dr_call, kTasks, 1, v + 2);
                p = Xi<double>{ v[0], v[1], v[2], v[3] };
            }
            init_state(s, seed, 0ull, tick);                                   // :226-229
            beyond = -1;
            tick_next = tick + 1;
            if (noisy) {
                double o[4];
                add_obs_noise(s, noise_std, seed, 0ull, tick, 1u, o);
                ctrl->obs[0] = o[0]; ctrl->obs[1] = o[1]; ctrl->obs[2] = o[2]; ctrl->obs[3] = o[3];
            }
        } else if (op == kOpSetState) {
            s = State<double>{ ld_sys_f64(&ctrl->arg[0]), ld_sys_f64(&ctrl->arg[1]), ld_sys_f64(&ctrl->arg[2]), ld_sys_f64(&ctrl->arg[3]) };
            beyond = (int32_t)ld_sys_u64(&ctrl->arg_u64[0]);
        } else if (op == kOpSetXi) {
            p = Xi<double>{ ld_sys_f64(&ctrl->arg[0]), ld_sys_f64(&ctrl->arg[1]), ld_sys_f64(&ctrl->arg[2]), ld_sys_f64(&ctrl->arg[3]) };
        } else if (op == kOpConfig) {
            // arg_u64: [0] seed, [1] euler, [2] dr_type, [3] noisy, [4] step clock, [5] look-ahead; arg: [0] noise std,
            // [1..4] a, [5..8] b, [9..12] floor
            seed = ld_sys_u64(&ctrl->arg_u64[0]);
            euler = ld_sys_u64(&ctrl->arg_u64[1]) != 0;
            noisy = ld_sys_u64(&ctrl->arg_u64[3]) != 0;
            tick_next = ld_sys_u64(&ctrl->arg_u64[4]);
            lookahead = ld_sys_u64(&ctrl->arg_u64[5]) != 0;
            noise_std = ld_sys_f64(&ctrl->arg[0]);
            dr.dr_type = (int)ld_sys_u64(&ctrl->arg_u64[2]); dr.dim = 4;
            for (int k = 0; k < 4; ++k) {
                dr.a[k] = ld_sys_f64(&ctrl->arg[1 + k]); dr.b[k] = ld_sys_f64(&ctrl->arg[5 + k]);
                dr.floor[k] = ld_sys_f64(&ctrl->arg[9 + k]);
            }
        }
        if (!is_step) { one_valid = false; two_valid = false; published = false; nz_valid = false; nz_after_valid = false; }
        if (!ahead) {
            scalar_put_state(ctrl->state, s);
            if (!is_step) {
                ctrl->xi[0] = p.gravity; ctrl->xi[1] = p.cart_mass; ctrl->xi[2] = p.pole_mass; ctrl->xi[3] = p.pole_length;
                ctrl->violations = viol;
                ctrl->beyond = beyond;
            }
            // release at system scope: the results above are visible to the host before the acknowledgement
            asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(&ctrl->ack), "r"(seq) : "memory");
        }
        last = seq;
        idle_since = global_timer_ns();
        if (op == kOpExit) break;
    }
    save->state[0] = s.x; save->state[1] = s.x_dot; save->state[2] = s.theta; save->state[3] = s.theta_dot;
    save->xi[0] = p.gravity; save->xi[1] = p.cart_mass; save->xi[2] = p.pole_mass; save->xi[3] = p.pole_length;
    save->noise_std = noise_std; save->beyond = beyond; save->euler = euler ? 1 : 0; save->noisy = noisy ? 1 : 0; save->seed = seed;
    save->lookahead = lookahead ? 1 : 0;
    save->tick_next = tick_next;
    save->last_seq = last; save->dr = dr; save->valid = 1u;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&ctrl->exited) = lease_id;
}

}  // namespace renv
