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
// The kernel is a LEASE, not a daemon: after `lease_ns` without a request it saves its registers to `save` (device
// memory) and exits, setting ctrl->exited = lease id; the host relaunches it with the next request (the request word
// stays pending in host memory and is served by the new instance).  So a device-wide synchronise elsewhere in the
// process stalls for at most one lease, and nothing can spin forever on either side.
#pragma once
#include "../../include/renv.h"
#include "renv_kernels.cuh"

namespace renv {

enum ScalarOp : uint32_t { kOpStep0 = 0, kOpStep1 = 1, kOpReset = 2, kOpSetState = 3, kOpSetXi = 4, kOpConfig = 5,
                           kOpExit = 6 };

// What survives between kernel instances (device memory, 512 bytes, zero-initialised by the caller once).
struct ScalarSave {
    double state[4], xi[4], noise_std;
    int32_t beyond, euler, noisy;
    uint32_t last_seq, valid;
    uint64_t seed;
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
    uint64_t seed = save->seed;
    uint32_t last = save->valid ? save->last_seq : ld_sys_u32(&ctrl->ack);
    DrCfg4<double> dr = save->dr;
    if (!save->valid) dr.dr_type = kDrNone;

    uint64_t idle_since = global_timer_ns();
    for (;;) {
        // While the host is busy between two calls, BOTH possible next steps are computed (action 0 and action 1):
        // when the request arrives the answer is a select, and the ~1 us of dependent fp64 arithmetic of a single
        // thread is off the round trip.
        State<double> next[2] = { s, s };
        bool term[2];
        {
            const Derived<double> d = derive(p);
            term[0] = dynamics(next[0], p, d, 0, euler);
            term[1] = dynamics(next[1], p, d, 1, euler);
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
        if (op > kOpStep1) __threadfence_system();      // the arguments were written before the request word
        if (op <= kOpStep1) {
            // RandomCartPoleEnv.step, no auto-reset, TimeLimit left to the gym wrapper (as in the reference stack)
            s = next[op];
            const bool terminated = term[op];
            double reward = 1.0;                                               // :207-212
            if (terminated) {                                                  // :213-222
                reward = beyond < 0 ? 1.0 : 0.0;
                beyond = beyond < 0 ? 0 : beyond + 1;
            }
            // reward, done and steps_beyond_done travel as one 16-byte store
            int4 tail;
            tail.x = (int)(__double_as_longlong(reward) & 0xffffffffll); tail.y = (int)(__double_as_longlong(reward) >> 32);
            tail.z = terminated ? 1 : 0; tail.w = beyond;
            *reinterpret_cast<int4 *>(&ctrl->reward) = tail;
            if (noisy) {                        // Noisy variant: obs = state + std N(0, I); the host wrote this step's tick
                double o[4];
                add_obs_noise(s, noise_std, seed, 0ull, ld_sys_u64(&ctrl->arg_u64[0]), 0u, o);
                ctrl->obs[0] = o[0]; ctrl->obs[1] = o[1]; ctrl->obs[2] = o[2]; ctrl->obs[3] = o[3];
            }
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
            // arg_u64: [0] seed, [1] euler, [2] dr_type, [3] noisy; arg: [0] noise std, [1..4] a, [5..8] b, [9..12] floor
            seed = ld_sys_u64(&ctrl->arg_u64[0]);
            euler = ld_sys_u64(&ctrl->arg_u64[1]) != 0;
            noisy = ld_sys_u64(&ctrl->arg_u64[3]) != 0;
            noise_std = ld_sys_f64(&ctrl->arg[0]);
            dr.dr_type = (int)ld_sys_u64(&ctrl->arg_u64[2]); dr.dim = 4;
            for (int k = 0; k < 4; ++k) {
                dr.a[k] = ld_sys_f64(&ctrl->arg[1 + k]); dr.b[k] = ld_sys_f64(&ctrl->arg[5 + k]);
                dr.floor[k] = ld_sys_f64(&ctrl->arg[9 + k]);
            }
        }
        *reinterpret_cast<double2 *>(&ctrl->state[0]) = make_double2(s.x, s.x_dot);
        *reinterpret_cast<double2 *>(&ctrl->state[2]) = make_double2(s.theta, s.theta_dot);
        if (op > kOpStep1) {
            ctrl->xi[0] = p.gravity; ctrl->xi[1] = p.cart_mass; ctrl->xi[2] = p.pole_mass; ctrl->xi[3] = p.pole_length;
            ctrl->violations = viol;
            ctrl->beyond = beyond;
        }
        // release at system scope: the results above are visible to the host before the acknowledgement
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(&ctrl->ack), "r"(seq) : "memory");
        last = seq;
        idle_since = global_timer_ns();
        if (op == kOpExit) break;
    }
    save->state[0] = s.x; save->state[1] = s.x_dot; save->state[2] = s.theta; save->state[3] = s.theta_dot;
    save->xi[0] = p.gravity; save->xi[1] = p.cart_mass; save->xi[2] = p.pole_mass; save->xi[3] = p.pole_length;
    save->noise_std = noise_std; save->beyond = beyond; save->euler = euler ? 1 : 0; save->noisy = noisy ? 1 : 0; save->seed = seed;
    save->last_seq = last; save->dr = dr; save->valid = 1u;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(&ctrl->exited) = lease_id;
}

}  // namespace renv
