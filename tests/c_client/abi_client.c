/* A C program that uses librenv_b200.so exactly as include/renv.h describes -- no Python, no torch: cudaMalloc'd
 * buffers, the caller's stream, plain structs.  It runs reset + K steps (fp64, uniform DR, TimeLimit, auto-reset) and
 * a fused rollout, and checks every done flag / counter / state against the C oracle (oracle/cartpole_oracle.c,
 * linked in as the checker).  Built and run by tests/test_gpu_c_client.py:
 *
 *   nvcc -o build/abi_client tests/c_client/abi_client.c oracle/cartpole_oracle.c -Iinclude \
 *        -Lrandom_envs_b200 -lrenv_b200 -Xlinker -rpath=$PWD/random_envs_b200 -lm
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "renv.h"

void oracle_closed_loop_f64(int64_t n, double *state, double *xi, int32_t *elapsed, uint32_t *episode, uint64_t seed,
                            uint64_t env_id0, uint64_t tick0, int K, int max_steps, int euler, const uint8_t *actions,
                            const double *w, double b, const double *lo, const double *hi, double *stats,
                            uint8_t *done_log, uint8_t *trunc_log, double *state_log);
void oracle_init_state_f64(uint64_t seed, uint64_t id, uint64_t ep, double s[4]);
void oracle_xi_uniform_f64(uint64_t seed, uint64_t id, uint64_t ep, uint32_t purpose, int dim, const double *lo,
                           const double *hi, double *out);
void oracle_random_actions(int64_t n, uint64_t env_id0, uint64_t seed, uint64_t step, uint8_t *out);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define RV(x) do { int r_ = (x); if (r_ != RENV_OK) { fprintf(stderr, "%s -> %d (%s)\n", #x, r_, renv_strerror(r_)); return 3; } } while (0)

int main(void)
{
    enum { N = 4099, LD = 4100, K = 60, MAXSTEPS = 25 };
    const uint64_t seed = 77, id0 = 1000;
    const double lo[4] = { 2.0, 0.5, 0.05, 0.1 }, hi[4] = { 20.0, 3.0, 0.3, 1.0 };
    if (renv_abi_version() != RENV_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }

    double *d_state, *d_xi, *d_reward, *d_stats;
    int32_t *d_elapsed, *d_beyond; uint32_t *d_episode;
    uint8_t *d_action, *d_done, *d_trunc;
    unsigned long long *d_viol;
    CK(cudaMalloc((void **)&d_state, 4 * LD * 8)); CK(cudaMalloc((void **)&d_xi, 4 * LD * 8));
    CK(cudaMalloc((void **)&d_reward, LD * 8)); CK(cudaMalloc((void **)&d_stats, 6 * 8));
    CK(cudaMalloc((void **)&d_elapsed, LD * 4)); CK(cudaMalloc((void **)&d_beyond, LD * 4));
    CK(cudaMalloc((void **)&d_episode, LD * 4)); CK(cudaMalloc((void **)&d_viol, 8 * RENV_NUM_COUNTERS));
    CK(cudaMalloc((void **)&d_action, LD)); CK(cudaMalloc((void **)&d_done, LD)); CK(cudaMalloc((void **)&d_trunc, LD));
    CK(cudaMemset(d_state, 0, 4 * LD * 8)); CK(cudaMemset(d_xi, 0, 4 * LD * 8)); CK(cudaMemset(d_elapsed, 0, LD * 4));
    CK(cudaMemset(d_episode, 0, LD * 4)); CK(cudaMemset(d_beyond, 0xff, LD * 4)); CK(cudaMemset(d_viol, 0, 8 * RENV_NUM_COUNTERS));
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));

    renv_cartpole_env env;
    env.state = d_state; env.xi = d_xi; env.elapsed = d_elapsed; env.episode = d_episode; env.beyond = d_beyond;
    env.elapsed16 = NULL; env.progress = NULL;
    env.n = N; env.ld = LD; env.env_id0 = id0; env.seed = seed;
    renv_dr_cfg dr;
    memset(&dr, 0, sizeof dr);
    dr.dr_type = RENV_DR_UNIFORM; dr.dim = 4;
    for (int k = 0; k < 4; ++k) { dr.a[k] = lo[k]; dr.b[k] = hi[k]; dr.lb[k] = 0.1; }

    /* reset at tick 0, then K steps at ticks 1..K with the library's own random actions */
    RV(renv_cartpole_reset_f64(&env, NULL, 0, &dr, d_viol, stream));
    static double h_state[4 * LD], h_xi[4 * LD], r_state[4 * N], r_xi[4 * N];
    static int32_t r_el[N]; static uint32_t r_ep[N];
    static uint8_t h_done[LD], h_trunc[LD], acts[K * N], r_done[K * N], r_trunc[K * N];
    static double r_log[(size_t)K * 4 * N];
    CK(cudaMemcpyAsync(h_state, d_state, sizeof h_state, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_xi, d_xi, sizeof h_xi, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int i = 0; i < N; ++i) {
        double s0[4], x0[4];
        oracle_init_state_f64(seed, id0 + i, 0, s0);
        oracle_xi_uniform_f64(seed, id0 + i, 0, 1, 4, lo, hi, x0);
        for (int c = 0; c < 4; ++c) {
            if (h_state[c * LD + i] != s0[c] || h_xi[4 * i + c] != x0[c]) { fprintf(stderr, "reset mismatch env %d\n", i); return 4; }
            r_state[c * N + i] = s0[c]; r_xi[c * N + i] = x0[c];
        }
        r_el[i] = 0; r_ep[i] = 1;
    }
    for (int k = 0; k < K; ++k) oracle_random_actions(N, id0, seed, 1 + k, acts + (size_t)k * N);
    double r_stats[6] = { 0, 0, 0, INFINITY, -INFINITY, 0 };
    oracle_closed_loop_f64(N, r_state, r_xi, r_el, r_ep, seed, id0, 1, K, MAXSTEPS, 1, acts, NULL, 0.0, lo, hi, r_stats,
                           r_done, r_trunc, r_log);
    double worst = 0.0;
    long dones = 0;
    for (int k = 0; k < K; ++k) {
        RV(renv_random_actions_u8(d_action, N, id0, seed, (uint32_t)(1 + k), stream));
        RV(renv_cartpole_step_f64(&env, d_action, d_reward, d_done, d_trunc, RENV_EULER, MAXSTEPS, 1, 1 + k, &dr, d_viol, stream));
        CK(cudaMemcpyAsync(h_state, d_state, sizeof h_state, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(h_done, d_done, LD, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(h_trunc, d_trunc, LD, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (int i = 0; i < N; ++i) {
            if (h_done[i] != r_done[(size_t)k * N + i] || h_trunc[i] != r_trunc[(size_t)k * N + i]) {
                fprintf(stderr, "flag mismatch step %d env %d\n", k, i); return 5;
            }
            dones += h_done[i];
            for (int c = 0; c < 4; ++c) {
                const double e = fabs(h_state[c * LD + i] - r_log[((size_t)k * 4 + c) * N + i]);
                if (e > worst) worst = e;
            }
        }
    }
    if (!(worst <= 1e-9)) { fprintf(stderr, "state error %g\n", worst); return 6; }

    /* fused rollout: 40 more steps under a linear policy; the statistics vector must equal the oracle's */
    const double w[4] = { 0.0, 0.0, 1.0, 0.0 };
    double h_stats[6] = { 0, 0, 0, INFINITY, -INFINITY, 0 }, o_stats[6] = { 0, 0, 0, INFINITY, -INFINITY, 0 };
    CK(cudaMemcpyAsync(d_stats, h_stats, sizeof h_stats, cudaMemcpyHostToDevice, stream));
    RV(renv_cartpole_rollout_f64(&env, w, 0.0, 40, RENV_EULER, MAXSTEPS, 1 + K, &dr, d_stats, d_viol, stream));
    CK(cudaMemcpyAsync(h_stats, d_stats, sizeof h_stats, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    oracle_closed_loop_f64(N, r_state, r_xi, r_el, r_ep, seed, id0, 1 + K, 40, MAXSTEPS, 1, NULL, w, 0.0, lo, hi, o_stats,
                           NULL, NULL, NULL);
    for (int q = 0; q < 6; ++q)
        if (h_stats[q] != o_stats[q]) { fprintf(stderr, "stats[%d] %g != %g\n", q, h_stats[q], o_stats[q]); return 7; }

    /* argument contract: errors are return codes, never crashes */
    if (renv_cartpole_step_f64(NULL, d_action, d_reward, d_done, d_trunc, 0, 0, 1, 0, NULL, NULL, stream) != RENV_E_NULL) return 8;
    env.ld = N - 1;
    if (renv_cartpole_reset_f64(&env, NULL, 0, NULL, NULL, stream) != RENV_E_SIZE) return 9;
    env.ld = LD; dr.dr_type = 9;
    if (renv_cartpole_reset_f64(&env, NULL, 0, &dr, NULL, stream) != RENV_E_DRTYPE) return 10;

    printf("abi_client ok: %d envs x %d steps, %ld episode ends, max |state - oracle| = %.3g, rollout stats equal\n", N, K,
           dones, worst);
    return 0;
}
