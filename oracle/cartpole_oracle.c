/* Plain-C restatement of the RandomCartPole-v0 hot path.   TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this
 * (see oracle/__init__.py).  It exists because the pure-Python port cannot reach the
 * sizes the GPU parity tests need (1e6..1e7 env-steps) in seconds.
 *
 * Parity pin: CPython floats are C doubles, `math.sin/cos` are libm sin/cos and
 * `x ** 2` is libm pow(x, 2.0), so compiled with -ffp-contract=off -fno-builtin this file
 * performs the SAME IEEE operations in the SAME order as the reference's
 * random_envs/random_cartpole.py:176-205.  tests/test_oracle_cartpole.py checks it
 * bit-for-bit against the golden trajectories produced by the real reference.
 *
 * Sections:
 *   1. dynamics + termination          random_cartpole.py:176-205
 *   2. Philox4x32-10 and the draw spec (NOT from the reference: it restates the framework's
 *      own published RNG contract, DESIGN.md "RNG contract", so that closed-loop runs with
 *      auto-reset can be replayed on the CPU; Philox itself is Salmon et al., SC'11)
 *   3. closed loop: TimeLimit (gym 0.21) + SyncVectorEnv auto-reset + uniform DR resample
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* ---- 1. dynamics ----------------------------------------------------------------------- */
#define FORCE_MAG 10.0
#define POLEMASS_LENGTH 0.05   /* 0.1*0.5 frozen in __init__ (random_cartpole.py:79); set_task never refreshes it */
#define TAU 0.02
#define X_THRESHOLD 2.4

static double theta_threshold(void) { return 12 * 2 * 3.141592653589793 / 360; }

/* returns terminated flag; s = {x, x_dot, theta, theta_dot}; xi = {g, m_cart, m_pole, l} */
static int step_one(double *s, const double *xi, int action, int euler)
{
    double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    double gravity = xi[0], cart_mass = xi[1], pole_mass = xi[2], pole_length = xi[3];
    double total_mass = pole_mass + cart_mass;                      /* :166 */
    double force = action == 1 ? FORCE_MAG : -FORCE_MAG;            /* :177 */
    double costheta = cos(theta), sintheta = sin(theta);            /* :178-179 */
    double temp = (force + POLEMASS_LENGTH * pow(theta_dot, 2.0) * sintheta) / total_mass;      /* :183 */
    double thetaacc = (gravity * sintheta - costheta * temp) /
        (pole_length * (4.0 / 3.0 - pole_mass * pow(costheta, 2.0) / total_mass));              /* :184 */
    double xacc = temp - POLEMASS_LENGTH * thetaacc * costheta / total_mass;                    /* :185 */
    if (euler) {                                                    /* :187-191 */
        x = x + TAU * x_dot;
        x_dot = x_dot + TAU * xacc;
        theta = theta + TAU * theta_dot;
        theta_dot = theta_dot + TAU * thetaacc;
    } else {                                                        /* :192-196 */
        x_dot = x_dot + TAU * xacc;
        x = x + TAU * x_dot;
        theta_dot = theta_dot + TAU * thetaacc;
        theta = theta + TAU * theta_dot;
    }
    s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;
    double thr = theta_threshold();
    return (x < -X_THRESHOLD) || (x > X_THRESHOLD) || (theta < -thr) || (theta > thr);          /* :200-205 */
}

/* state, xi: SoA (4, n) row-major.  One step for every env, no wrappers. */
void oracle_step_batch(int64_t n, double *state, const double *xi, const uint8_t *action,
                       int euler, uint8_t *terminated)
{
    for (int64_t i = 0; i < n; ++i) {
        double s[4] = { state[i], state[n + i], state[2 * n + i], state[3 * n + i] };
        double p[4] = { xi[i], xi[n + i], xi[2 * n + i], xi[3 * n + i] };
        terminated[i] = (uint8_t)step_one(s, p, action[i], euler);
        state[i] = s[0]; state[n + i] = s[1]; state[2 * n + i] = s[2]; state[3 * n + i] = s[3];
    }
}

/* ---- 2. Philox4x32-10 + draw spec -------------------------------------------------------- */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { PURPOSE_INIT = 0, PURPOSE_XI = 1, PURPOSE_ACTION = 2, PURPOSE_TASKS = 3 };

/* counter = (id[31:0], id[47:32] | tick[47:32] << 16, tick[31:0], purpose << 24 | slot); tick = the env's step clock
 * at the launch that started the episode (DESIGN.md "RNG contract"). */
static void draw(uint64_t seed, uint64_t id, uint64_t tick, uint32_t purpose, uint32_t sub, uint32_t r[4])
{
    uint32_t c1 = ((uint32_t)(id >> 32) & 0xffffu) | (((uint32_t)(tick >> 32) & 0xffffu) << 16);
    uint32_t ctr[4] = { (uint32_t)id, c1, (uint32_t)tick, (purpose << 24) | sub };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    oracle_philox4x32_10(ctr, key, r);
}

static double u01_f64(uint32_t hi, uint32_t lo)   /* numpy's 53-bit recipe */
{
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
static float u01_f32(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

/* k-th U[0,1) of stream (seed,id,episode,purpose) at attempt t -- fp64 packing: 2 per call */
static double uniform_f64(uint64_t seed, uint64_t id, uint64_t ep, uint32_t purpose, int t, int k)
{
    uint32_t r[4];
    draw(seed, id, ep, purpose, (uint32_t)(t * 16 + k / 2), r);
    return (k & 1) ? u01_f64(r[2], r[3]) : u01_f64(r[0], r[1]);
}
static float uniform_f32(uint64_t seed, uint64_t id, uint64_t ep, uint32_t purpose, int t, int k)
{
    uint32_t r[4];
    draw(seed, id, ep, purpose, (uint32_t)(t * 16 + k / 4), r);
    return u01_f32(r[k & 3]);
}

/* reset draws: s0 ~ U(-0.05, 0.05)^4  (random_cartpole.py:227; numpy: low + (high-low)*u) */
void oracle_init_state_f64(uint64_t seed, uint64_t id, uint64_t ep, double s[4])
{
    for (int k = 0; k < 4; ++k) s[k] = -0.05 + 0.1 * uniform_f64(seed, id, ep, PURPOSE_INIT, 0, k);
}
void oracle_init_state_f32(uint64_t seed, uint64_t id, uint64_t ep, float s[4])
{
    for (int k = 0; k < 4; ++k) s[k] = fmaf(0.1f, uniform_f32(seed, id, ep, PURPOSE_INIT, 0, k), -0.05f);
}
/* uniform DR draw of one xi vector (random_env.py:151), attempt 0 only */
void oracle_xi_uniform_f64(uint64_t seed, uint64_t id, uint64_t ep, uint32_t purpose, int dim,
                           const double *lo, const double *hi, double *out)
{
    for (int k = 0; k < dim; ++k) out[k] = lo[k] + (hi[k] - lo[k]) * uniform_f64(seed, id, ep, purpose, 0, k);
}
void oracle_xi_uniform_f32(uint64_t seed, uint64_t id, uint64_t ep, uint32_t purpose, int dim,
                           const double *lo, const double *hi, float *out)
{
    for (int k = 0; k < dim; ++k) {
        float l = (float)lo[k], h = (float)hi[k];
        out[k] = fmaf(h - l, uniform_f32(seed, id, ep, purpose, 0, k), l);
    }
}
/* raw uniforms, for checking the truncnorm / gaussian transforms against scipy on the host */
void oracle_uniforms_f64(uint64_t seed, uint64_t id, uint64_t ep, uint32_t purpose, int t, int dim, double *out)
{
    for (int k = 0; k < dim; ++k) out[k] = uniform_f64(seed, id, ep, purpose, t, k);
}
void oracle_uniforms_f32(uint64_t seed, uint64_t id, uint64_t ep, uint32_t purpose, int t, int dim, float *out)
{
    for (int k = 0; k < dim; ++k) out[k] = uniform_f32(seed, id, ep, purpose, t, k);
}

/* Bernoulli(1/2) actions: env e at step t uses bit (t & 127) of ITS 128-bit block (e, t >> 7): one Philox block
 * carries an env's actions for 128 consecutive steps (a fused rollout draws one block per 128 env-steps) */
void oracle_random_actions(int64_t n, uint64_t env_id0, uint64_t seed, uint64_t step, uint8_t *out)
{
    step &= 0xffffffffull;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t e = env_id0 + (uint64_t)i;
        uint32_t r[4];
        draw(seed, e, step >> 7, PURPOSE_ACTION, 0, r);
        out[i] = (uint8_t)((r[(step >> 5) & 3] >> (step & 31)) & 1u);
    }
}

/* ---- 3. closed loop ------------------------------------------------------------------------ */
/* One env's worth of: [policy] -> step -> TimeLimit -> auto-reset (+ uniform DR resample).
 * state/xi SoA (4,n); elapsed (n) int32; episode (n) uint32 = episodes started (statistics);
 * step k of the loop runs at clock tick0 + k: that tick keys the Philox draws of a reset happening in it.
 * actions: (K,n) uint8 when w == NULL, else ignored and a = [w.s + b > 0] (left-to-right sum).
 * lo/hi NULL => xi untouched on reset.  max_steps <= 0 => no TimeLimit.
 * stats[6] += {episodes, sum R, sum R^2, min R, max R, sum length}  (min/max start at +/-inf by caller)
 * done_log (K,n), trunc_log (K,n), state_log (K,4,n) optional (NULL to skip); state_log holds the
 * state returned by step k, i.e. AFTER auto-reset, as the vector env does.
 */
void oracle_closed_loop_f64(int64_t n, double *state, double *xi, int32_t *elapsed, uint32_t *episode,
                            uint64_t seed, uint64_t env_id0, uint64_t tick0, int K, int max_steps, int euler,
                            const uint8_t *actions, const double *w, double b,
                            const double *lo, const double *hi,
                            double *stats, uint8_t *done_log, uint8_t *trunc_log, double *state_log)
{
    for (int64_t i = 0; i < n; ++i) {
        double s[4] = { state[i], state[n + i], state[2 * n + i], state[3 * n + i] };
        double p[4] = { xi[i], xi[n + i], xi[2 * n + i], xi[3 * n + i] };
        int32_t el = elapsed[i];
        uint32_t ep = episode[i];
        uint64_t id = env_id0 + (uint64_t)i;
        for (int k = 0; k < K; ++k) {
            int a;
            if (w) {
                double acc = w[0] * s[0];
                acc = acc + w[1] * s[1];
                acc = acc + w[2] * s[2];
                acc = acc + w[3] * s[3];
                acc = acc + b;
                a = acc > 0.0;
            } else {
                a = actions[(int64_t)k * n + i];
            }
            int done = step_one(s, p, a, euler);
            int trunc = 0;
            el += 1;
            if (max_steps > 0 && el >= max_steps) { trunc = !done; done = 1; }
            if (done) {
                double ret = (double)el;   /* reward is 1.0 on every step incl. the terminal one (:207-212) */
                stats[0] += 1.0; stats[1] += ret; stats[2] += ret * ret;
                if (ret < stats[3]) stats[3] = ret;
                if (ret > stats[4]) stats[4] = ret;
                stats[5] += (double)el;
                ep += 1; el = 0;
                if (lo) oracle_xi_uniform_f64(seed, id, tick0 + (uint64_t)k, PURPOSE_XI, 4, lo, hi, p);
                oracle_init_state_f64(seed, id, tick0 + (uint64_t)k, s);
            }
            if (done_log) done_log[(int64_t)k * n + i] = (uint8_t)done;
            if (trunc_log) trunc_log[(int64_t)k * n + i] = (uint8_t)trunc;
            if (state_log) for (int c = 0; c < 4; ++c) state_log[((int64_t)k * 4 + c) * n + i] = s[c];
        }
        state[i] = s[0]; state[n + i] = s[1]; state[2 * n + i] = s[2]; state[3 * n + i] = s[3];
        xi[i] = p[0]; xi[n + i] = p[1]; xi[2 * n + i] = p[2]; xi[3 * n + i] = p[3];
        elapsed[i] = el; episode[i] = ep;
    }
}
