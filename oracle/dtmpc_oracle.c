/*
 * dtmpc_oracle.c -- CPU restatement of the reference's tube-MPC hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see dtmpc_oracle.h).  Dense, unbatched, scalar C
 * that follows /root/reference line by line; every function cites the lines
 * it restates.  Compile with -O2 -ffp-contract=off (no FMA contraction: the
 * reference evaluates 0-d torch ops one at a time).
 */
#include "dtmpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NX 4
#define NU 2

static int finite_n(const double *v, int n) {
    for (int i = 0; i < n; ++i)
        if (!isfinite(v[i])) return 0;
    return 1;
}

/* ---------------------------------------------------------------------- */
/* core/systems/dubins_obstacles.py:16-30  h_i = dx*dx + dy*dy - r**2      */
static double h_circle(const orc_env *env, int i, const double x[3]) {
    double dx = x[0] - env->obs_cx[i];
    double dy = x[1] - env->obs_cy[i];
    return dx * dx + dy * dy - (env->obs_r[i] * env->obs_r[i]);
}

/* core/systems/dubins_obstacles.py:41-69 (smooth-min), :95-106 (min),
 * :16-30 (single circle, tube_mpc.py:700-704); empty list -> 1 (:57-61). */
double orc_h(const orc_env *env, const double x[3]) {
    int M = env->n_obs;
    if (M == 0) return 1.0;
    if (env->obs_agg == ORC_AGG_SINGLE) return h_circle(env, 0, x);
    double H[ORC_MAX_OBS] = {0};
    for (int i = 0; i < M; ++i) H[i] = h_circle(env, i, x);
    if (env->obs_agg == ORC_AGG_MIN) {
        double m = H[0];
        for (int i = 1; i < M; ++i)
            if (H[i] < m) m = H[i];
        return m;
    }
    double z[ORC_MAX_OBS];
    double zmax = -INFINITY;
    for (int i = 0; i < M; ++i) {
        z[i] = (-env->obs_beta) * H[i];
        if (z[i] > zmax) zmax = z[i];
    }
    double s = 0.0;
    for (int i = 0; i < M; ++i) s += exp(z[i] - zmax);
    double lse = zmax + log(s);
    return -(1.0 / env->obs_beta) * lse;
}

/* core/systems/dubins_obstacles.py:33-38 (circle), :72-92 (softmax weights),
 * :109-117 (argmin obstacle); empty list -> 0 (:81-82). */
void orc_grad_h(const orc_env *env, const double x[3], double g[3]) {
    int M = env->n_obs;
    g[0] = g[1] = g[2] = 0.0;
    if (M == 0) return;
    if (env->obs_agg == ORC_AGG_SINGLE) {
        g[0] = 2.0 * (x[0] - env->obs_cx[0]);
        g[1] = 2.0 * (x[1] - env->obs_cy[0]);
        return;
    }
    double H[ORC_MAX_OBS];
    for (int i = 0; i < M; ++i) H[i] = h_circle(env, i, x);
    if (env->obs_agg == ORC_AGG_MIN) {
        int a = 0;
        for (int i = 1; i < M; ++i)
            if (H[i] < H[a]) a = i; /* torch.argmin: first minimum */
        g[0] = 2.0 * (x[0] - env->obs_cx[a]);
        g[1] = 2.0 * (x[1] - env->obs_cy[a]);
        return;
    }
    double z[ORC_MAX_OBS], w[ORC_MAX_OBS];
    double zmax = -INFINITY;
    for (int i = 0; i < M; ++i) {
        z[i] = (-env->obs_beta) * H[i];
        if (z[i] > zmax) zmax = z[i];
    }
    double s = 0.0;
    for (int i = 0; i < M; ++i) {
        w[i] = exp(z[i] - zmax);
        s += w[i];
    }
    for (int i = 0; i < M; ++i) {
        double wi = w[i] / s;
        g[0] += wi * (2.0 * (x[0] - env->obs_cx[i]));
        g[1] += wi * (2.0 * (x[1] - env->obs_cy[i]));
        g[2] += wi * 0.0;
    }
}

/* core/barrier.py:36-59 */
double orc_B(double z, double alpha, double eps) {
    double a = alpha > eps ? alpha : eps; /* torch.maximum */
    if (z >= a) {
        double zc = z < eps ? eps : z;
        return 1.0 / zc;
    }
    double d = z - a;
    return (1.0 / a) - d / (a * a) + (d * d) / (a * a * a);
}

/* core/systems/dubins_aug_jac.py:26-40 */
double orc_dB(double z, double alpha, double eps) {
    double a = alpha > eps ? alpha : eps;
    if (z >= a) {
        double zc = z < eps ? eps : z;
        return -1.0 / (zc * zc);
    }
    double d = z - a;
    return -(1.0 / (a * a)) + (2.0 * d) / (a * a * a);
}

/* core/systems/dubins.py:24-43 */
void orc_dubins_step(const orc_env *env, const double x[3], const double u[2], double xn[3]) {
    double dt = env->dt;
    xn[0] = x[0] + dt * u[0] * cos(x[2]);
    xn[1] = x[1] + dt * u[0] * sin(x[2]);
    xn[2] = x[2] + dt * u[1];
}

/* core/barrier.py:75-108 through the closure core/tube_mpc.py:268-279 / :816-821
 * (h_nom = h - tight, core/tube_mpc.py:152-153, 273-274). */
void orc_fhat(const orc_env *env, const orc_ocp *p, const double xh[4], const double u[2], double out[4]) {
    double xn[3];
    orc_dubins_step(env, xh, u, xn);
    double h_next = orc_h(env, xn) - p->tight;
    double h_curr = orc_h(env, xh) - p->tight;
    double B_next = orc_B(h_next, p->alpha, p->eps_dyn);
    double B_curr = orc_B(h_curr, p->alpha, p->eps_dyn);
    out[0] = xn[0];
    out[1] = xn[1];
    out[2] = xn[2];
    out[3] = B_next - p->gamma * (B_curr - xh[3]);
}

/* core/barrier.py:111-120 (initialisation uses the YAML eps: tube_mpc.py:135-160, 707-712) */
double orc_b0(const orc_env *env, const orc_ocp *p, const double x[3]) {
    return orc_B(orc_h(env, x) - p->tight, p->alpha, p->eps_jac);
}

/* core/systems/dubins_aug_jac.py:61-139 (dense 4x4 / 4x2, row-major) */
void orc_aug_jac(const orc_env *env, const orc_ocp *p, const double xh[4], const double u[2], double A[16],
                 double Bm[8]) {
    double dt = env->dt;
    double th = xh[2], v = u[0];
    double c = cos(th), s = sin(th);
    double A3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double B3[6] = {0, 0, 0, 0, 0, 0};
    A3[0 * 3 + 2] = -dt * v * s;
    A3[1 * 3 + 2] = dt * v * c;
    B3[0 * 2 + 0] = dt * c;
    B3[1 * 2 + 0] = dt * s;
    B3[2 * 2 + 1] = dt;

    double xn[3];
    xn[0] = xh[0] + dt * u[0] * cos(xh[2]);
    xn[1] = xh[1] + dt * u[0] * sin(xh[2]);
    xn[2] = xh[2] + dt * u[1];

    double off = p->tight_in_jac ? p->tight : 0.0;
    double h_curr = orc_h(env, xh) - off;
    double h_next = orc_h(env, xn) - off;
    double dh_curr[3], dh_next[3];
    orc_grad_h(env, xh, dh_curr);
    orc_grad_h(env, xn, dh_next);
    double dB_curr = orc_dB(h_curr, p->alpha, p->eps_jac);
    double dB_next = orc_dB(h_next, p->alpha, p->eps_jac);

    double gn[3], gc[3];
    for (int i = 0; i < 3; ++i) {
        gn[i] = dB_next * dh_next[i];
        gc[i] = p->gamma * dB_curr * dh_curr[i];
    }
    double row_x[3], row_u[2];
    for (int j = 0; j < 3; ++j) {
        double acc = 0.0;
        for (int i = 0; i < 3; ++i) acc += gn[i] * A3[i * 3 + j];
        row_x[j] = acc - gc[j];
    }
    for (int j = 0; j < 2; ++j) {
        double acc = 0.0;
        for (int i = 0; i < 3; ++i) acc += gn[i] * B3[i * 2 + j];
        row_u[j] = acc;
    }
    memset(A, 0, 16 * sizeof(double));
    memset(Bm, 0, 8 * sizeof(double));
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) A[i * 4 + j] = A3[i * 3 + j];
        for (int j = 0; j < 2; ++j) Bm[i * 2 + j] = B3[i * 2 + j];
    }
    for (int j = 0; j < 3; ++j) A[3 * 4 + j] = row_x[j];
    A[3 * 4 + 3] = p->gamma;
    Bm[3 * 2 + 0] = row_u[0];
    Bm[3 * 2 + 1] = row_u[1];
}

/* core/control.py:61-64 */
void orc_clamp(const orc_env *env, const double u[2], double out[2]) {
    for (int j = 0; j < 2; ++j) {
        double v = u[j];
        if (v < env->u_min[j]) v = env->u_min[j];
        if (v > env->u_max[j]) v = env->u_max[j];
        out[j] = v;
    }
}

/* core/control.py:66-70 */
void orc_active(const orc_env *env, const double u[2], int act[2]) {
    for (int j = 0; j < 2; ++j)
        act[j] = (u[j] <= (env->u_min[j] + env->active_tol)) || (u[j] >= (env->u_max[j] - env->active_tol));
}

/* run_nominal.py:32-34 */
static double wrap_angle(double e) { return atan2(sin(e), cos(e)); }

/* core/tube_mpc.py:823-827, 875-880 (281-286, 358-364); run_nominal.py:297-302 */
double orc_stage_cost(const orc_ocp *p, const double xh[4], const double u[2], const double xref[3],
                      const double uref[2]) {
    double dx[3], du[2];
    for (int i = 0; i < 3; ++i) dx[i] = xh[i] - xref[i];
    if (p->angle_wrap) dx[2] = wrap_angle(dx[2]);
    for (int j = 0; j < 2; ++j) du[j] = u[j] - uref[j];
    double sx = 0.0, su = 0.0;
    for (int i = 0; i < 3; ++i) sx += p->Q[i] * dx[i] * dx[i];
    for (int j = 0; j < 2; ++j) su += p->R[j] * du[j] * du[j];
    return sx + su + p->qb * (xh[3] * xh[3]);
}

/* core/tube_mpc.py:829-832, 882-885 (288-293, 366-371); run_nominal.py:304-309 */
double orc_terminal_cost(const orc_ocp *p, const double xh[4], const double xref[3]) {
    double dx[3];
    for (int i = 0; i < 3; ++i) dx[i] = xh[i] - xref[i];
    if (p->angle_wrap) dx[2] = wrap_angle(dx[2]);
    double sx = 0.0;
    for (int i = 0; i < 3; ++i) sx += p->Qf[i] * dx[i] * dx[i];
    return sx + p->qb * (xh[3] * xh[3]);
}

/* core/cost_derivs.py:58-76, 110-129 (l_xx, l_uu diagonal, l_ux = 0);
 * angle-wrapped shifted target: run_nominal.py:311-315 */
static void stage_derivs(const orc_ocp *p, const double xh[4], const double u[2], const double xref[3],
                         const double uref[2], double lx[4], double lu[2], double lxx[16], double luu[4]) {
    double tgt[3] = {xref[0], xref[1], xref[2]};
    if (p->angle_wrap) tgt[2] = xh[2] - wrap_angle(xh[2] - xref[2]);
    for (int i = 0; i < 3; ++i) lx[i] = 2.0 * p->Q[i] * (xh[i] - tgt[i]);
    lx[3] = 2.0 * p->qb * xh[3];
    for (int j = 0; j < 2; ++j) lu[j] = 2.0 * p->R[j] * (u[j] - uref[j]);
    memset(lxx, 0, 16 * sizeof(double));
    memset(luu, 0, 4 * sizeof(double));
    for (int i = 0; i < 3; ++i) lxx[i * 4 + i] = 2.0 * p->Q[i];
    lxx[15] = 2.0 * p->qb;
    for (int j = 0; j < 2; ++j) luu[j * 2 + j] = 2.0 * p->R[j];
}

/* core/cost_derivs.py:132-146 + barrier overrides core/tube_mpc.py:298-303, 837-842;
 * run_nominal.py:317-324 adds instead of assigning (same value, Q13) */
static void terminal_derivs(const orc_ocp *p, const double xh[4], const double xref[3], double px[4],
                            double pxx[16]) {
    double tgt[3] = {xref[0], xref[1], xref[2]};
    if (p->angle_wrap) tgt[2] = xh[2] - wrap_angle(xh[2] - xref[2]);
    for (int i = 0; i < 3; ++i) px[i] = 2.0 * p->Qf[i] * (xh[i] - tgt[i]);
    px[3] = 2.0 * p->qb * xh[3];
    memset(pxx, 0, 16 * sizeof(double));
    for (int i = 0; i < 3; ++i) pxx[i * 4 + i] = 2.0 * p->Qf[i];
    pxx[15] = 2.0 * p->qb;
}

/* dense row-major helpers: C[m,n] = A[m,k] @ B[k,n]; sequential dot products */
static void mm(const double *A, const double *B, double *C, int m, int k, int n) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
            for (int l = 0; l < k; ++l) acc += A[i * k + l] * B[l * n + j];
            C[i * n + j] = acc;
        }
}
static void tr(const double *A, double *At, int m, int n) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}

/* torch.linalg.solve on a 2x2 (LAPACK getrf/getrs: partial-pivot LU), nrhs columns.
 * A row-major [2,2], Bm row-major [2,nrhs] -> X [2,nrhs]. core/ddp.py:243-244, 58 */
static void solve2(const double A[4], const double *Bm, double *X, int nrhs) {
    double a00 = A[0], a01 = A[1], a10 = A[2], a11 = A[3];
    int swap = fabs(a10) > fabs(a00);
    if (swap) {
        double t;
        t = a00; a00 = a10; a10 = t;
        t = a01; a01 = a11; a11 = t;
    }
    double l = a10 / a00;
    double u11 = a11 - l * a01;
    for (int c = 0; c < nrhs; ++c) {
        double b0 = Bm[(swap ? 1 : 0) * nrhs + c];
        double b1 = Bm[(swap ? 0 : 1) * nrhs + c];
        double y1 = b1 - l * b0;
        double x1 = y1 / u11;
        double x0 = (b0 - a01 * x1) / a00;
        X[0 * nrhs + c] = x0;
        X[1 * nrhs + c] = x1;
    }
}

/* core/ddp.py:23-60 */
static void solve_reduced(const double A[4], const double *Bm, const int act[2], double *X, int nrhs) {
    for (int i = 0; i < 2 * nrhs; ++i) X[i] = 0.0;
    if (act[0] && act[1]) return;
    if (!act[0] && !act[1]) {
        solve2(A, Bm, X, nrhs);
        return;
    }
    int f = act[0] ? 1 : 0;
    for (int c = 0; c < nrhs; ++c) X[f * nrhs + c] = Bm[f * nrhs + c] / A[f * 2 + f];
}

static void ref_at(const orc_ocp *p, const double *Xref, const double *Uref, int k, double xr[3], double ur[2]) {
    if (Xref) {
        xr[0] = Xref[k * 3 + 0]; xr[1] = Xref[k * 3 + 1]; xr[2] = Xref[k * 3 + 2];
    } else {
        xr[0] = p->target[0]; xr[1] = p->target[1]; xr[2] = p->target[2];
    }
    if (ur) {
        if (Uref) { ur[0] = Uref[k * 2 + 0]; ur[1] = Uref[k * 2 + 1]; }
        else { ur[0] = 0.0; ur[1] = 0.0; }
    }
}

/* core/ddp.py:102-307 */
int orc_ilqr_solve(const orc_env *env, const orc_ocp *p, const double x0[4], const double *V_init,
                   const double *Xref, const double *Uref, double *X, double *V, double *K, double *kff,
                   int *iters, int *alpha_idx, double *costs, double *margins) {
    const int T = p->T;
    double *A_seq = (double *)malloc(sizeof(double) * T * 16);
    double *B_seq = (double *)malloc(sizeof(double) * T * 8);
    double *lx = (double *)malloc(sizeof(double) * T * 4);
    double *lu = (double *)malloc(sizeof(double) * T * 2);
    double *lxx = (double *)malloc(sizeof(double) * T * 16);
    double *luu = (double *)malloc(sizeof(double) * T * 4);
    double *Xn = (double *)malloc(sizeof(double) * (T + 1) * 4);
    double *Vn = (double *)malloc(sizeof(double) * T * 2);
    double *Xb = (double *)malloc(sizeof(double) * (T + 1) * 4);
    double *Vb = (double *)malloc(sizeof(double) * T * 2);
    int status = ORC_OK;
    int it_done = 0;

    /* :127-131  V = clamp(V_init); X = rollout */
    for (int k = 0; k < T; ++k) orc_clamp(env, V_init + 2 * k, V + 2 * k);
    memcpy(X, x0, 4 * sizeof(double));
    for (int k = 0; k < T; ++k) orc_fhat(env, p, X + 4 * k, V + 2 * k, X + 4 * (k + 1));
    for (int i = 0; i < p->max_iter; ++i) {
        alpha_idx[i] = -1;
        costs[i] = NAN;
        margins[i] = NAN;
    }
    if (K) memset(K, 0, sizeof(double) * T * 8);
    if (kff) memset(kff, 0, sizeof(double) * T * 2);

    double prev_cost = 0.0;
    int have_prev = 0;
    double *Kl = (double *)malloc(sizeof(double) * T * 8);
    double *kl = (double *)malloc(sizeof(double) * T * 2);

    for (int it = 0; it < p->max_iter; ++it) {
        /* :172-205 linearisation along the trajectory */
        double cost = 0.0;
        for (int k = 0; k < T; ++k) {
            double xr[3], ur[2];
            ref_at(p, Xref, Uref, k, xr, ur);
            if (!finite_n(X + 4 * k, 4) || !finite_n(V + 2 * k, 2)) { status = ORC_NONFINITE; goto done; }
            orc_aug_jac(env, p, X + 4 * k, V + 2 * k, A_seq + 16 * k, B_seq + 8 * k);
            if (!finite_n(A_seq + 16 * k, 16) || !finite_n(B_seq + 8 * k, 8)) { status = ORC_NONFINITE; goto done; }
            stage_derivs(p, X + 4 * k, V + 2 * k, xr, ur, lx + 4 * k, lu + 2 * k, lxx + 16 * k, luu + 4 * k);
            if (!finite_n(lx + 4 * k, 4) || !finite_n(lu + 2 * k, 2)) { status = ORC_NONFINITE; goto done; }
            cost = cost + orc_stage_cost(p, X + 4 * k, V + 2 * k, xr, ur);
            if (!isfinite(cost)) { status = ORC_NONFINITE; goto done; }
        }
        double Vx[4], Vxx[16];
        {
            double xr[3];
            ref_at(p, Xref, Uref, T, xr, NULL);
            terminal_derivs(p, X + 4 * T, xr, Vx, Vxx);
            if (!finite_n(Vx, 4)) { status = ORC_NONFINITE; goto done; }
            cost = cost + orc_terminal_cost(p, X + 4 * T, xr);
            if (!isfinite(cost)) { status = ORC_NONFINITE; goto done; }
        }

        /* :208-254 backward pass */
        for (int k = T - 1; k >= 0; --k) {
            const double *A = A_seq + 16 * k;
            const double *Bk = B_seq + 8 * k;
            double At[16], Bt[8], t44[16], t24[8];
            tr(A, At, 4, 4);
            tr(Bk, Bt, 4, 2);
            double Qx[4], Qu[2], Qxx[16], Qux[8], Qxu[8], Quu[4], Quur[4];
            double tv[4];
            mm(At, Vx, tv, 4, 4, 1);
            for (int i = 0; i < 4; ++i) Qx[i] = lx[4 * k + i] + tv[i];
            mm(Bt, Vx, tv, 2, 4, 1);
            for (int i = 0; i < 2; ++i) Qu[i] = lu[2 * k + i] + tv[i];
            mm(At, Vxx, t44, 4, 4, 4);
            mm(t44, A, Qxx, 4, 4, 4);
            for (int i = 0; i < 16; ++i) Qxx[i] = lxx[16 * k + i] + Qxx[i];
            mm(Bt, Vxx, t24, 2, 4, 4);
            mm(t24, A, Qux, 2, 4, 4);
            for (int i = 0; i < 8; ++i) Qux[i] = 0.0 + Qux[i]; /* l_ux = 0 */
            tr(Qux, Qxu, 2, 4);
            mm(t24, Bk, Quu, 2, 4, 2);
            for (int i = 0; i < 4; ++i) Quu[i] = luu[4 * k + i] + Quu[i];
            if (!finite_n(Quu, 4)) { status = ORC_NONFINITE; goto done; }
            for (int i = 0; i < 4; ++i) Quur[i] = Quu[i];
            Quur[0] += p->reg;
            Quur[3] += p->reg;

            double Kk[8], kk[2];
            solve2(Quur, Qux, Kk, 4);
            solve2(Quur, Qu, kk, 1);
            for (int i = 0; i < 8; ++i) Kk[i] = -Kk[i];
            for (int i = 0; i < 2; ++i) kk[i] = -kk[i];
            if (!finite_n(Kk, 8) || !finite_n(kk, 2)) { status = ORC_NONFINITE; goto done; }
            memcpy(Kl + 8 * k, Kk, sizeof(Kk));
            memcpy(kl + 2 * k, kk, sizeof(kk));

            /* V_x = Q_x + K.T @ Q_uu @ kff + K.T @ Q_u + Q_xu @ kff */
            double Kt[8], KtQuu[8], t4a[4], t4b[4], t4c[4];
            tr(Kk, Kt, 2, 4);
            mm(Kt, Quu, KtQuu, 4, 2, 2);
            mm(KtQuu, kk, t4a, 4, 2, 1);
            mm(Kt, Qu, t4b, 4, 2, 1);
            mm(Qxu, kk, t4c, 4, 2, 1);
            double nVx[4], nVxx[16], m1[16], m2[16], m3[16];
            for (int i = 0; i < 4; ++i) nVx[i] = Qx[i] + t4a[i] + t4b[i] + t4c[i];
            /* V_xx = Q_xx + K.T @ Q_uu @ K + K.T @ Q_ux + Q_xu @ K */
            mm(KtQuu, Kk, m1, 4, 2, 4);
            mm(Kt, Qux, m2, 4, 2, 4);
            mm(Qxu, Kk, m3, 4, 2, 4);
            for (int i = 0; i < 16; ++i) nVxx[i] = Qxx[i] + m1[i] + m2[i] + m3[i];
            if (!finite_n(nVx, 4) || !finite_n(nVxx, 16)) { status = ORC_NONFINITE; goto done; }
            memcpy(Vx, nVx, sizeof(nVx));
            memcpy(Vxx, nVxx, sizeof(nVxx));
        }
        if (K) memcpy(K, Kl, sizeof(double) * T * 8);
        if (kff) memcpy(kff, kl, sizeof(double) * T * 2);

        /* :256-296 forward pass, every alpha evaluated, strict '<' */
        int have_best = 0, best_i = -1;
        double best_cost = 0.0, second = INFINITY;
        for (int ai = 0; ai < p->n_alpha; ++ai) {
            double alpha = p->alphas[ai];
            double x[4];
            memcpy(x, x0, sizeof(x));
            memcpy(Xn, x0, 4 * sizeof(double));
            for (int k = 0; k < T; ++k) {
                double dx[4], du[2], u[2], xnext[4];
                for (int i = 0; i < 4; ++i) dx[i] = x[i] - X[4 * k + i];
                for (int j = 0; j < 2; ++j) {
                    double acc = 0.0;
                    for (int i = 0; i < 4; ++i) acc += Kl[8 * k + 4 * j + i] * dx[i];
                    du[j] = kl[2 * k + j] + acc;
                }
                for (int j = 0; j < 2; ++j) u[j] = V[2 * k + j] + alpha * du[j];
                orc_clamp(env, u, u);
                Vn[2 * k] = u[0];
                Vn[2 * k + 1] = u[1];
                orc_fhat(env, p, x, u, xnext);
                memcpy(x, xnext, sizeof(x));
                memcpy(Xn + 4 * (k + 1), x, sizeof(x));
            }
            if (!finite_n(Xn, 4 * (T + 1)) || !finite_n(Vn, 2 * T)) { status = ORC_NONFINITE; goto done; }
            double cc = 0.0;
            for (int k = 0; k < T; ++k) {
                double xr[3], ur[2];
                ref_at(p, Xref, Uref, k, xr, ur);
                cc = cc + orc_stage_cost(p, Xn + 4 * k, Vn + 2 * k, xr, ur);
            }
            {
                double xr[3];
                ref_at(p, Xref, Uref, T, xr, NULL);
                cc = cc + orc_terminal_cost(p, Xn + 4 * T, xr);
            }
            if (!isfinite(cc)) { status = ORC_NONFINITE; goto done; }
            if (!have_best || cc < best_cost) {
                if (have_best && best_cost < second) second = best_cost;
                have_best = 1;
                best_cost = cc;
                best_i = ai;
                memcpy(Xb, Xn, sizeof(double) * 4 * (T + 1));
                memcpy(Vb, Vn, sizeof(double) * 2 * T);
            } else if (cc < second) {
                second = cc;
            }
        }
        if (!have_best) { status = ORC_NO_CANDIDATE; goto done; }
        memcpy(X, Xb, sizeof(double) * 4 * (T + 1));
        memcpy(V, Vb, sizeof(double) * 2 * T);
        alpha_idx[it] = best_i;
        costs[it] = best_cost;
        margins[it] = second - best_cost;
        it_done = it + 1;
        /* :303-305 */
        if (have_prev && fabs(prev_cost - best_cost) < p->tol) break;
        prev_cost = best_cost;
        have_prev = 1;
    }
done:
    *iters = it_done;
    free(A_seq); free(B_seq); free(lx); free(lu); free(lxx); free(luu);
    free(Xn); free(Vn); free(Xb); free(Vb); free(Kl); free(kl);
    return status;
}

/* core/ddp.py:317-427 */
int orc_sensitivity(const orc_env *env, const orc_ocp *p, const double *X, const double *V, const double *gx,
                    const double *gu, double *dX, double *dV, double *dL, int *active) {
    const int T = p->T;
    double *A_seq = (double *)malloc(sizeof(double) * T * 16);
    double *B_seq = (double *)malloc(sizeof(double) * T * 8);
    double *Ks = (double *)malloc(sizeof(double) * T * 8);
    double *ks = (double *)malloc(sizeof(double) * T * 2);
    double *Vxx_seq = (double *)malloc(sizeof(double) * (T + 1) * 16);
    double *tVx_seq = (double *)malloc(sizeof(double) * (T + 1) * 4);
    int status = ORC_OK;

    /* :345-359 linearise at the final (X, V); hessians are constant diagonals */
    for (int k = 0; k < T; ++k) orc_aug_jac(env, p, X + 4 * k, V + 2 * k, A_seq + 16 * k, B_seq + 8 * k);
    double lxx[16], luu[4], pxx[16];
    memset(lxx, 0, sizeof(lxx));
    memset(luu, 0, sizeof(luu));
    memset(pxx, 0, sizeof(pxx));
    for (int i = 0; i < 3; ++i) { lxx[i * 4 + i] = 2.0 * p->Q[i]; pxx[i * 4 + i] = 2.0 * p->Qf[i]; }
    lxx[15] = 2.0 * p->qb;
    pxx[15] = 2.0 * p->qb;
    for (int j = 0; j < 2; ++j) luu[j * 2 + j] = 2.0 * p->R[j];

    double Vxx[16], tVx[4];
    memcpy(Vxx, pxx, sizeof(Vxx));
    memcpy(tVx, gx + 4 * T, sizeof(tVx));
    memcpy(Vxx_seq + 16 * T, Vxx, sizeof(Vxx));
    memcpy(tVx_seq + 4 * T, tVx, sizeof(tVx));

    /* :375-410 backward */
    for (int k = T - 1; k >= 0; --k) {
        const double *A = A_seq + 16 * k;
        const double *Bk = B_seq + 8 * k;
        double At[16], Bt[8], t44[16], t24[8];
        tr(A, At, 4, 4);
        tr(Bk, Bt, 4, 2);
        double Qxx[16], Qxu[8], Qux[8], Quu[4], Quur[4];
        mm(At, Vxx, t44, 4, 4, 4);
        mm(t44, A, Qxx, 4, 4, 4);
        for (int i = 0; i < 16; ++i) Qxx[i] = lxx[i] + Qxx[i];
        mm(t44, Bk, Qxu, 4, 4, 2); /* l_ux.T = 0 */
        mm(Bt, Vxx, t24, 2, 4, 4);
        mm(t24, A, Qux, 2, 4, 4);
        mm(t24, Bk, Quu, 2, 4, 2);
        for (int i = 0; i < 4; ++i) Quu[i] = luu[i] + Quu[i];
        double tQu[2], tQx[4], tv[4];
        mm(Bt, tVx, tv, 2, 4, 1);
        for (int i = 0; i < 2; ++i) tQu[i] = gu[2 * k + i] + tv[i];
        mm(At, tVx, tv, 4, 4, 1);
        for (int i = 0; i < 4; ++i) tQx[i] = gx[4 * k + i] + tv[i];
        for (int i = 0; i < 4; ++i) Quur[i] = Quu[i];
        Quur[0] += 1e-9; /* :373 */
        Quur[3] += 1e-9;
        int act[2];
        orc_active(env, V + 2 * k, act);
        active[2 * k] = act[0];
        active[2 * k + 1] = act[1];
        double Kk[8], kk[2];
        solve_reduced(Quur, Qux, act, Kk, 4);
        solve_reduced(Quur, tQu, act, kk, 1);
        for (int i = 0; i < 8; ++i) Kk[i] = -Kk[i];
        for (int i = 0; i < 2; ++i) kk[i] = -kk[i];
        memcpy(Ks + 8 * k, Kk, sizeof(Kk));
        memcpy(ks + 2 * k, kk, sizeof(kk));
        double t4[4], m[16];
        mm(Qxu, kk, t4, 4, 2, 1);
        for (int i = 0; i < 4; ++i) tVx[i] = tQx[i] + t4[i];
        mm(Qxu, Kk, m, 4, 2, 4);
        for (int i = 0; i < 16; ++i) Vxx[i] = Qxx[i] + m[i];
        memcpy(Vxx_seq + 16 * k, Vxx, sizeof(Vxx));
        memcpy(tVx_seq + 4 * k, tVx, sizeof(tVx));
    }

    /* :412-425 forward */
    memset(dX, 0, sizeof(double) * 4 * (T + 1));
    memset(dV, 0, sizeof(double) * 2 * T);
    memset(dL, 0, sizeof(double) * 4 * (T + 1));
    for (int k = 0; k < T; ++k) {
        double t2[2], t4a[4], t4b[4];
        mm(Ks + 8 * k, dX + 4 * k, t2, 2, 4, 1);
        for (int j = 0; j < 2; ++j) dV[2 * k + j] = ks[2 * k + j] + t2[j];
        for (int j = 0; j < 2; ++j)
            if (active[2 * k + j]) dV[2 * k + j] = 0.0;
        mm(A_seq + 16 * k, dX + 4 * k, t4a, 4, 4, 1);
        mm(B_seq + 8 * k, dV + 2 * k, t4b, 4, 2, 1);
        for (int i = 0; i < 4; ++i) dX[4 * (k + 1) + i] = t4a[i] + t4b[i];
        mm(Vxx_seq + 16 * k, dX + 4 * k, t4a, 4, 4, 1);
        for (int i = 0; i < 4; ++i) dL[4 * k + i] = tVx_seq[4 * k + i] + t4a[i];
    }
    {
        double t4a[4];
        mm(Vxx_seq + 16 * T, dX + 4 * T, t4a, 4, 4, 1);
        for (int i = 0; i < 4; ++i) dL[4 * T + i] = tVx_seq[4 * T + i] + t4a[i];
    }
    if (!finite_n(dX, 4 * (T + 1)) || !finite_n(dV, 2 * T) || !finite_n(dL, 4 * (T + 1))) status = ORC_NONFINITE;
    free(A_seq); free(B_seq); free(Ks); free(ks); free(Vxx_seq); free(tVx_seq);
    return status;
}

/* d B_alpha / d alpha_eff  (derivative of core/barrier.py:56-58 in the relaxed branch) */
static double dB_dalpha(double z, double alpha, double eps) {
    double a = alpha > eps ? alpha : eps;
    double w = alpha > eps ? 1.0 : (alpha == eps ? 0.5 : 0.0); /* torch.maximum backward */
    if (z >= a) return 0.0;
    double d = z - a;
    return w * (-3.0 * d * d / (a * a * a * a));
}

/* core/ift.py:35-92 with the gradient-time closures core/tube_mpc.py:464-487, 562-584,
 * differentiated by hand w.r.t. the constrained parameter values. */
void orc_ift_accumulate(const orc_env *env, const orc_ocp *p, const double *X, const double *V, const double *Xref,
                        const double *Uref, const double *dX, const double *dV, const double *dL, double g12[12],
                        double *gXref, double *gUref) {
    const int T = p->T;
    double gQ[3] = {0, 0, 0}, gR[2] = {0, 0}, gQf[3] = {0, 0, 0}, gqb = 0, ga = 0, gg = 0, gt = 0;
    for (int k = 0; k < T; ++k) {
        double xr[3], ur[2];
        ref_at(p, Xref, Uref, k, xr, ur);
        const double *x = X + 4 * k, *u = V + 2 * k;
        for (int i = 0; i < 3; ++i) gQ[i] += 2.0 * (x[i] - xr[i]) * dX[4 * k + i];
        for (int j = 0; j < 2; ++j) gR[j] += 2.0 * (u[j] - ur[j]) * dV[2 * k + j];
        gqb += 2.0 * x[3] * dX[4 * k + 3];
        if (gXref)
            for (int i = 0; i < 3; ++i) gXref[3 * k + i] = -2.0 * p->Q[i] * dX[4 * k + i];
        if (gUref)
            for (int j = 0; j < 2; ++j) gUref[2 * k + j] = -2.0 * p->R[j] * dV[2 * k + j];
        /* dynamics term: delta_lambda_{k+1}[3] * b_next(theta) */
        double lam = dL[4 * (k + 1) + 3];
        double xn[3];
        orc_dubins_step(env, x, u, xn);
        double hn = orc_h(env, xn) - p->tight;
        double hc = orc_h(env, x) - p->tight;
        double Bc = orc_B(hc, p->alpha, p->eps_jac);
        ga += lam * (dB_dalpha(hn, p->alpha, p->eps_jac) - p->gamma * dB_dalpha(hc, p->alpha, p->eps_jac));
        gg += lam * (-(Bc - x[3]));
        gt += lam * (-orc_dB(hn, p->alpha, p->eps_jac) + p->gamma * orc_dB(hc, p->alpha, p->eps_jac));
    }
    {
        double xr[3];
        ref_at(p, Xref, Uref, T, xr, NULL);
        const double *x = X + 4 * T;
        for (int i = 0; i < 3; ++i) gQf[i] = 2.0 * (x[i] - xr[i]) * dX[4 * T + i];
        gqb += 2.0 * x[3] * dX[4 * T + 3];
        if (gXref)
            for (int i = 0; i < 3; ++i) gXref[3 * T + i] = -2.0 * p->Qf[i] * dX[4 * T + i];
    }
    g12[0] = gQ[0]; g12[1] = gQ[1]; g12[2] = gQ[2];
    g12[3] = gR[0]; g12[4] = gR[1];
    g12[5] = gQf[0]; g12[6] = gQf[1]; g12[7] = gQf[2];
    g12[8] = gqb; g12[9] = ga; g12[10] = gg; g12[11] = gt;
}

static void load_theta(orc_ocp *o, const double *th) {
    o->Q[0] = th[0]; o->Q[1] = th[1]; o->Q[2] = th[2];
    o->R[0] = th[3]; o->R[1] = th[4];
    o->Qf[0] = th[5]; o->Qf[1] = th[6]; o->Qf[2] = th[7];
    o->qb = th[8]; o->alpha = th[9]; o->gamma = th[10]; o->tight = th[11];
}

/* The library itself is single-threaded; callers parallelise over instance
 * slices (oracle/dtmpc_oracle.py uses a thread pool; ctypes drops the GIL). */
int orc_num_threads(void) { return 1; }

/* One closed-loop step's worth of solves per instance:
 * core/tube_mpc.py:844-976 (paper mode) / :305-599 (full mode). */
int orc_tube_solve_batch(const orc_env *env, const orc_ocp *nom_cfg, const orc_ocp *aux_cfg, long B,
                         const double *x0n, const double *x0a, const double *Vn0, const double *Va0,
                         const double *theta_n, long theta_n_stride, const double *theta_a, long theta_a_stride,
                         int mode, int n_threads, orc_tube_out *o) {
    const int T = nom_cfg->T;
    const int MI_n = nom_cfg->max_iter, MI_a = aux_cfg->max_iter;
    int any_bad = 0;
    (void)n_threads;
    for (long i = 0; i < B; ++i) {
        orc_ocp pn = *nom_cfg, pa = *aux_cfg;
        load_theta(&pn, theta_n + i * theta_n_stride);
        load_theta(&pa, theta_a + i * theta_a_stride);
        double *Xn = o->Xn + i * (T + 1) * 4, *Un = o->Un + i * T * 2;
        double *Xa = o->Xa + i * (T + 1) * 4, *Ua = o->Ua + i * T * 2;
        double *Xref = (double *)malloc(sizeof(double) * (T + 1) * 3);
        double *gx = (double *)malloc(sizeof(double) * (T + 1) * 4);
        double *gu = (double *)calloc((size_t)T * 2, sizeof(double));
        int *act = (int *)malloc(sizeof(int) * T * 2);
        double *gXr = (double *)malloc(sizeof(double) * (T + 1) * 3);
        double *gUr = (double *)malloc(sizeof(double) * T * 2);
        int st = 0;
        st |= orc_ilqr_solve(env, &pn, x0n + 4 * i, Vn0 + i * T * 2, NULL, NULL, Xn, Un,
                             o->Kn ? o->Kn + i * T * 8 : NULL, o->kn ? o->kn + i * T * 2 : NULL, &o->iters_n[i],
                             o->aidx_n + i * MI_n, o->costs_n + i * MI_n, o->marg_n + i * MI_n);
        for (int k = 0; k <= T; ++k)
            for (int j = 0; j < 3; ++j) Xref[3 * k + j] = Xn[4 * k + j];
        if (!st)
            st |= orc_ilqr_solve(env, &pa, x0a + 4 * i, Va0 + i * T * 2, Xref, Un, Xa, Ua,
                                 o->Ka ? o->Ka + i * T * 8 : NULL, o->ka ? o->ka + i * T * 2 : NULL,
                                 &o->iters_a[i], o->aidx_a + i * MI_a, o->costs_a + i * MI_a,
                                 o->marg_a + i * MI_a) << 4;
        if (!st) {
            /* upper loss core/tube_mpc.py:916-919 and its gradients :932-944 */
            double lt = 0.0, lb = 0.0;
            for (int k = 0; k <= T; ++k) {
                for (int j = 0; j < 3; ++j) {
                    double d = Xa[4 * k + j] - Xn[4 * k + j];
                    lt += d * d;
                    gx[4 * k + j] = 2.0 * d;
                }
                lb += Xa[4 * k + 3] * Xa[4 * k + 3];
                gx[4 * k + 3] = 2.0 * Xa[4 * k + 3];
            }
            o->loss[i] = lt + lb;
            st |= orc_sensitivity(env, &pa, Xa, Ua, gx, gu, o->dXa + i * (T + 1) * 4, o->dVa + i * T * 2,
                                  o->dLa + i * (T + 1) * 4, act) << 8;
            orc_ift_accumulate(env, &pa, Xa, Ua, Xref, Un, o->dXa + i * (T + 1) * 4, o->dVa + i * T * 2,
                               o->dLa + i * (T + 1) * 4, o->g_aux + 12 * i, gXr, gUr);
            if (o->gXref) memcpy(o->gXref + i * (T + 1) * 3, gXr, sizeof(double) * (T + 1) * 3);
            if (o->gUref) memcpy(o->gUref + i * T * 2, gUr, sizeof(double) * T * 2);
            if (mode == 1) {
                /* core/tube_mpc.py:520-559: upper grads [g_xref, 0] and g_uref */
                for (int k = 0; k <= T; ++k) {
                    for (int j = 0; j < 3; ++j) gx[4 * k + j] = gXr[3 * k + j];
                    gx[4 * k + 3] = 0.0;
                }
                st |= orc_sensitivity(env, &pn, Xn, Un, gx, gUr, o->dXn + i * (T + 1) * 4, o->dVn + i * T * 2,
                                      o->dLn + i * (T + 1) * 4, act) << 12;
                orc_ift_accumulate(env, &pn, Xn, Un, NULL, NULL, o->dXn + i * (T + 1) * 4, o->dVn + i * T * 2,
                                   o->dLn + i * (T + 1) * 4, o->g_nom + 12 * i, NULL, NULL);
            }
        }
        o->status[i] = st;
        any_bad |= (st != 0);
        free(Xref); free(gx); free(gu); free(act); free(gXr); free(gUr);
    }
    return any_bad;
}
