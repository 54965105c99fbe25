// dtmpc_sens.cuh -- DDP-structured sensitivity sweep (reference core/ddp.py:317-427) fused with the
// closed-form IFT accumulation (reference core/ift.py:35-92; paper-mode core/tube_mpc.py:965-976).
//
// One problem per thread, static mapping (every instance does exactly one backward and one forward sweep
// of T steps, so there is no divergence to hide).  The backward sweep stores, per time step, what the
// forward sweep needs (gains, value blocks, linearisation, h) in a tile-major scratch: 32 consecutive instances form a
// tile whose rows are contiguous, ws[((i/32)*rows + row)*32 + i%32] -- a warp access is one 256 B segment and every
// address is the thread's base pointer plus an immediate (a batch-sized row stride costs 64-bit multiplies per access).
#pragma once
#include "dtmpc_math.cuh"

namespace dtmpc {

template <typename R> struct SensArgs {
    const R *X, *V, *theta, *target, *Xref, *Uref, *gx, *gu, *Xbar;
    long long theta_stride, target_stride, xref_row, gx_row;
    R *dX, *dV, *dL, *g_theta, *gXref, *gUref, *loss;
    int *status;
};

template <typename R> struct SensParams {
    EnvT<R> env;
    OcpT<R> ocp;
    SensArgs<R> io;
    long long B;
    R *ws;
};

constexpr int SENS_BLOCK = 128;
constexpr int SENS_ROWS = 8 + 2 + 4 + 16 + 8 + 1;  // K, k, tilde V_x, V_xx, linearisation, h  (per time step)

__host__ __device__ inline long long sens_ws_rows(int T) { return (long long)SENS_ROWS * T + 1; /* + h(x_T) */ }

// d total / d theta contributions of one stage (k < T) -- SURVEY App. A.10, derived from core/ift.py:62-85
template <typename R> struct GradAcc {
    R gQ0 = 0, gQ1 = 0, gQ2 = 0, gR0 = 0, gR1 = 0, gQf0 = 0, gQf1 = 0, gQf2 = 0, gqb = 0, ga = 0, gg = 0, gt = 0;
};

// torch.maximum(alpha, eps) passes the gradient to alpha when alpha > eps, half at a tie (core/barrier.py:49)
template <typename R> __device__ __forceinline__ R alpha_grad_weight(const Theta<R> &t) {
    return t.alpha > t.eps_jac ? R(1) : (t.alpha == t.eps_jac ? R(0.5) : R(0));
}

// B(z), dB/da (z) of the gradient-time barrier (a = a_jac, eps = eps_jac; core/barrier.py:36-59 and its derivative with
// respect to a in the relaxed branch) without data-dependent branches or divisions: c1 = 1/a, c4 = 1/a^4 are per-thread
// constants, 1/a^2 and 1/a^3 come with Theta.
template <typename R> struct JacConsts { R c1, c4; };
template <typename R> __device__ __forceinline__ JacConsts<R> jac_consts(const Theta<R> &t) {
    return JacConsts<R>{M<R>::rcp_(t.a_jac), t.c2_jac * t.c2_jac};
}
template <typename R> __device__ __forceinline__ R barrier_B_jac(const Theta<R> &t, const JacConsts<R> &c, R z) {
    using m = M<R>;
    const R safe = m::rcp_(m::max_(z, t.eps_jac));
    const R d = z - t.a_jac;
    const R relaxed = m::fma_(d * d, t.c3_jac, m::fma_(-d, t.c2_jac, c.c1));
    return z >= t.a_jac ? safe : relaxed;
}
template <typename R> __device__ __forceinline__ R barrier_dB_da_jac(const Theta<R> &t, const JacConsts<R> &c, R z) {
    const R d = z - t.a_jac;
    return z >= t.a_jac ? R(0) : (R(-3) * d * d) * c.c4;
}

template <typename R>
__device__ __forceinline__ void accumulate_stage(const Theta<R> &t, const JacConsts<R> &jc, GradAcc<R> &g, R aw, R px, R py,
                                                 R pth, R pb, R u0, R u1, R r0, R r1, R r2, R q0, R q1, R dx0, R dx1, R dx2,
                                                 R dx3, R du0, R du1, R lam_b_next, R h_curr, R h_next) {
    using m = M<R>;
    g.gQ0 = m::fma_(R(2) * (px - r0), dx0, g.gQ0);
    g.gQ1 = m::fma_(R(2) * (py - r1), dx1, g.gQ1);
    g.gQ2 = m::fma_(R(2) * (pth - r2), dx2, g.gQ2);
    g.gR0 = m::fma_(R(2) * (u0 - q0), du0, g.gR0);
    g.gR1 = m::fma_(R(2) * (u1 - q1), du1, g.gR1);
    g.gqb = m::fma_(R(2) * pb, dx3, g.gqb);
    // dynamics term  delta_lambda_{k+1}[b] * b_next(theta)  with  b_next = B(hn) - gamma (B(hc) - b)
    const R hn = h_next - t.tight, hc = h_curr - t.tight;
    const R Bc = barrier_B_jac(t, jc, hc);
    g.ga = m::fma_(lam_b_next, aw * (barrier_dB_da_jac(t, jc, hn) - t.gamma * barrier_dB_da_jac(t, jc, hc)), g.ga);
    g.gg = m::fma_(lam_b_next, -(Bc - pb), g.gg);
    g.gt = m::fma_(lam_b_next, t.gamma * barrier_dB_jac(t, hc) - barrier_dB_jac(t, hn), g.gt);
}

template <typename R> __device__ __forceinline__ void store_grads(R *out, const GradAcc<R> &g) {
    out[0] = g.gQ0; out[1] = g.gQ1; out[2] = g.gQ2;
    out[3] = g.gR0; out[4] = g.gR1;
    out[5] = g.gQf0; out[6] = g.gQf1; out[7] = g.gQf2;
    out[8] = g.gqb; out[9] = g.ga; out[10] = g.gg; out[11] = g.gt;
}

// NOBS > 0: obstacle count of a smooth-min environment fixed at compile time (as in the solver kernels)
#ifndef DTMPC_SENS_MAXNREG
#define DTMPC_SENS_MAXNREG 255
#endif
#ifndef DTMPC_SENS_MAXNREG_F32
#define DTMPC_SENS_MAXNREG_F32 255
#endif
template <typename R, int NOBS = -1>
__global__ void __maxnreg__(sizeof(R) == 8 ? DTMPC_SENS_MAXNREG : DTMPC_SENS_MAXNREG_F32) sens_kernel(const __grid_constant__ SensParams<R> P) {
    using m = M<R>;
    const long long i = (long long)blockIdx.x * SENS_BLOCK + threadIdx.x;
    if (i >= P.B) return;
    const EnvT<R> &E = P.env;
    const SensArgs<R> &a = P.io;
    const int T = P.ocp.T;
    R *const ws = P.ws + (i >> 5) * (sens_ws_rows(T) * 32) + (i & 31);
#define WS(row) ws[(row) * 32]
#define ROW(k, r) ((k) * SENS_ROWS + (r))
    Theta<R> th;
    load_theta(th, a.theta + i * a.theta_stride, P.ocp);
    const R *X = a.X + i * (4LL * (T + 1));
    const R *V = a.V + i * (2LL * T);
    const int gxr = (int)a.gx_row;  // 4, or 3: no barrier-state component
    const R *gxp = a.gx ? a.gx + i * ((long long)gxr * (T + 1)) : nullptr;
    const R *gup = a.gu ? a.gu + i * (2LL * T) : nullptr;
    const R *Xb = a.Xbar ? a.Xbar + i * (4LL * (T + 1)) : nullptr;
    const long long xrow = a.xref_row;
    const R *xrp = a.Xref ? a.Xref + i * (xrow * (T + 1)) : nullptr;
    const R *urp = a.Uref ? a.Uref + i * (2LL * T) : nullptr;
    R *gxo = a.gXref ? a.gXref + i * (3LL * (T + 1)) : nullptr;
    R *guo = a.gUref ? a.gUref + i * (2LL * T) : nullptr;
    const R gam = th.gamma, dt = E.dt;
    const JacConsts<R> jc = jac_consts(th);

    // upper-level gradient at step k: caller-supplied, or the tracking loss 2(x - xbar), 2b (core/tube_mpc.py:932-944)
    auto upper_gx = [&](int k, R px, R py, R pt, R pb, R &g0, R &g1, R &g2, R &g3) {
        if (gxp) {
            g0 = gxp[gxr * k]; g1 = gxp[gxr * k + 1]; g2 = gxp[gxr * k + 2]; g3 = gxr == 4 ? gxp[4 * k + 3] : R(0);
        } else {
            R b0, b1, b2, b3;
            ld4(Xb + 4 * k, b0, b1, b2, b3);
            g0 = R(2) * (px - b0);
            g1 = R(2) * (py - b1);
            g2 = R(2) * (pt - b2);
            g3 = R(2) * pb;
        }
    };

    // ---- backward sweep (core/ddp.py:361-410) -----------------------------
    R Vx0, Vx1, Vx2, Vx3;
    R V00, V01, V02, V03, V10, V11, V12, V13, V20, V21, V22, V23, V30, V31, V32, V33;
    R gnx, gny;
    R loss_t = R(0), loss_b = R(0);
    {
        R px, py, pt, pb;
        ld4(X + 4 * T, px, py, pt, pb);
        upper_gx(T, px, py, pt, pb, Vx0, Vx1, Vx2, Vx3);
        V00 = R(2) * th.Qf0; V11 = R(2) * th.Qf1; V22 = R(2) * th.Qf2; V33 = R(2) * th.qb;
        V01 = V02 = V03 = V10 = V12 = V13 = V20 = V21 = V23 = V30 = V31 = V32 = R(0);
        R hT;
        barrier_grad<R, NOBS>(E, th, px, py, gnx, gny, hT);
        WS(ROW(T, 0)) = hT;
    }
    const R lxx0 = R(2) * th.Q0, lxx1 = R(2) * th.Q1, lxx2 = R(2) * th.Q2, lxx3 = R(2) * th.qb;
    const R luu0 = R(2) * th.R0, luu1 = R(2) * th.R1;
    const R sreg = R(1e-9);  // core/ddp.py:373
    // one running pointer per stream (scratch rows of step k, state / control rows of the instance): every access is
    // pointer + immediate; indexing by k cost ~4 integer instructions of 64-bit address arithmetic per access
    R *wk = ws + (long long)T * (SENS_ROWS * 32);
    const R *Xk = X + 4LL * T, *Vk = V + 2LL * T;
    for (int k = T - 1; k >= 0; --k) {
        wk -= SENS_ROWS * 32;
        Xk -= 4;
        Vk -= 2;
#define WK(r) wk[(r) * 32]
        R px, py, pt, pb, u0, u1;
        ld4(Xk, px, py, pt, pb);
        ld2(Vk, u0, u1);
        R gcx, gcy, hk;
        barrier_grad<R, NOBS>(E, th, px, py, gcx, gcy, hk);
        Lin<R> L;
        linearize(E, th, pt, u0, gnx, gny, gcx, gcy, L);
        gnx = gcx;
        gny = gcy;
        R g0, g1, g2, g3;
        upper_gx(k, px, py, pt, pb, g0, g1, g2, g3);
        R gu0 = R(0), gu1 = R(0);
        if (gup) ld2(gup + 2 * k, gu0, gu1);

#define ROW_M(i)                                                                              \
    const R M##i##0 = m::fma_(V##i##3, L.r0, V##i##0);                                        \
    const R M##i##1 = m::fma_(V##i##3, L.r1, V##i##1);                                        \
    const R M##i##2 = m::fma_(V##i##3, L.r2, m::fma_(V##i##1, L.a12, m::fma_(V##i##0, L.a02, V##i##2))); \
    const R M##i##3 = V##i##3 * gam;
        ROW_M(0) ROW_M(1) ROW_M(2) ROW_M(3)
#undef ROW_M
#define COL_Q(j)                                                                              \
    R Qxx0##j = m::fma_(L.r0, M3##j, M0##j);                                                  \
    R Qxx1##j = m::fma_(L.r1, M3##j, M1##j);                                                  \
    R Qxx2##j = m::fma_(L.r2, M3##j, m::fma_(L.a12, M1##j, m::fma_(L.a02, M0##j, M2##j)));    \
    R Qxx3##j = gam * M3##j;
        COL_Q(0) COL_Q(1) COL_Q(2) COL_Q(3)
#undef COL_Q
        Qxx00 += lxx0; Qxx11 += lxx1; Qxx22 += lxx2; Qxx33 += lxx3;
#define COL_U(j)                                                                              \
    const R Qux0##j = m::fma_(L.s0, M3##j, m::fma_(L.b10, M1##j, L.b00 * M0##j));             \
    const R Qux1##j = dt * M2##j;
        COL_U(0) COL_U(1) COL_U(2) COL_U(3)
#undef COL_U
        // V_xx B columns (w, y); Q_xu = A^T V_xx B computed independently of Q_ux (core/ddp.py:380-381)
        const R w0 = m::fma_(V03, L.s0, m::fma_(V01, L.b10, V00 * L.b00));
        const R w1 = m::fma_(V13, L.s0, m::fma_(V11, L.b10, V10 * L.b00));
        const R w2 = m::fma_(V23, L.s0, m::fma_(V21, L.b10, V20 * L.b00));
        const R w3 = m::fma_(V33, L.s0, m::fma_(V31, L.b10, V30 * L.b00));
        const R y0 = V02 * dt, y1 = V12 * dt, y2 = V22 * dt, y3 = V32 * dt;
        const R Qxu00 = m::fma_(L.r0, w3, w0), Qxu01 = m::fma_(L.r0, y3, y0);
        const R Qxu10 = m::fma_(L.r1, w3, w1), Qxu11 = m::fma_(L.r1, y3, y1);
        const R Qxu20 = m::fma_(L.r2, w3, m::fma_(L.a12, w1, m::fma_(L.a02, w0, w2)));
        const R Qxu21 = m::fma_(L.r2, y3, m::fma_(L.a12, y1, m::fma_(L.a02, y0, y2)));
        const R Qxu30 = gam * w3, Qxu31 = gam * y3;
        const R Quu00 = luu0 + m::fma_(L.s0, w3, m::fma_(L.b10, w1, L.b00 * w0));
        const R Quu01 = m::fma_(L.s0, y3, m::fma_(L.b10, y1, L.b00 * y0));
        const R Quu10 = dt * w2;
        const R Quu11 = luu1 + dt * y2;
        const R tQu0 = gu0 + m::fma_(L.s0, Vx3, m::fma_(L.b10, Vx1, L.b00 * Vx0));
        const R tQu1 = gu1 + dt * Vx2;
        const R tQx0 = g0 + m::fma_(L.r0, Vx3, Vx0);
        const R tQx1 = g1 + m::fma_(L.r1, Vx3, Vx1);
        const R tQx2 = g2 + m::fma_(L.r2, Vx3, m::fma_(L.a12, Vx1, m::fma_(L.a02, Vx0, Vx2)));
        const R tQx3 = g3 + gam * Vx3;

        // active-set reduced solves (core/ddp.py:23-60, 394-401)
        bool a0, a1;
        active_u(E, u0, u1, a0, a1);
        R K00 = 0, K01 = 0, K02 = 0, K03 = 0, K10 = 0, K11 = 0, K12 = 0, K13 = 0, k0 = 0, k1 = 0;
        const R A00 = Quu00 + sreg, A11 = Quu11 + sreg;
        if (!a0 && !a1) {
            LU2<R> F;
            lu2_factor(F, A00, Quu01, Quu10, A11);
            lu2_solve(F, Qux00, Qux10, K00, K10);
            lu2_solve(F, Qux01, Qux11, K01, K11);
            lu2_solve(F, Qux02, Qux12, K02, K12);
            lu2_solve(F, Qux03, Qux13, K03, K13);
            lu2_solve(F, tQu0, tQu1, k0, k1);
        } else if (!a0) {
            const R inv = R(1) / A00;
            K00 = Qux00 * inv; K01 = Qux01 * inv; K02 = Qux02 * inv; K03 = Qux03 * inv;
            k0 = tQu0 * inv;
        } else if (!a1) {
            const R inv = R(1) / A11;
            K10 = Qux10 * inv; K11 = Qux11 * inv; K12 = Qux12 * inv; K13 = Qux13 * inv;
            k1 = tQu1 * inv;
        }
        K00 = -K00; K01 = -K01; K02 = -K02; K03 = -K03;
        K10 = -K10; K11 = -K11; K12 = -K12; K13 = -K13;
        k0 = -k0; k1 = -k1;
        // tilde V_x = tilde Q_x + Q_xu k ;  V_xx = Q_xx + Q_xu K   (core/ddp.py:403-404)
        Vx0 = tQx0 + m::fma_(Qxu01, k1, Qxu00 * k0);
        Vx1 = tQx1 + m::fma_(Qxu11, k1, Qxu10 * k0);
        Vx2 = tQx2 + m::fma_(Qxu21, k1, Qxu20 * k0);
        Vx3 = tQx3 + m::fma_(Qxu31, k1, Qxu30 * k0);
#define ENT_V(i, j) V##i##j = Qxx##i##j + m::fma_(Qxu##i##1, K1##j, Qxu##i##0 * K0##j);
        ENT_V(0, 0) ENT_V(0, 1) ENT_V(0, 2) ENT_V(0, 3)
        ENT_V(1, 0) ENT_V(1, 1) ENT_V(1, 2) ENT_V(1, 3)
        ENT_V(2, 0) ENT_V(2, 1) ENT_V(2, 2) ENT_V(2, 3)
        ENT_V(3, 0) ENT_V(3, 1) ENT_V(3, 2) ENT_V(3, 3)
#undef ENT_V
        // stash for the forward sweep (row 0: h, rows 1..38 below)
        st_stream(wk + (1) * 32, K00); st_stream(wk + (2) * 32, K01); st_stream(wk + (3) * 32, K02); st_stream(wk + (4) * 32, K03);
        st_stream(wk + (5) * 32, K10); st_stream(wk + (6) * 32, K11); st_stream(wk + (7) * 32, K12); st_stream(wk + (8) * 32, K13);
        st_stream(wk + (9) * 32, k0); st_stream(wk + (10) * 32, k1);
        st_stream(wk + (11) * 32, Vx0); st_stream(wk + (12) * 32, Vx1); st_stream(wk + (13) * 32, Vx2); st_stream(wk + (14) * 32, Vx3);
        st_stream(wk + (15) * 32, V00); st_stream(wk + (16) * 32, V01); st_stream(wk + (17) * 32, V02); st_stream(wk + (18) * 32, V03);
        st_stream(wk + (19) * 32, V10); st_stream(wk + (20) * 32, V11); st_stream(wk + (21) * 32, V12); st_stream(wk + (22) * 32, V13);
        st_stream(wk + (23) * 32, V20); st_stream(wk + (24) * 32, V21); st_stream(wk + (25) * 32, V22); st_stream(wk + (26) * 32, V23);
        st_stream(wk + (27) * 32, V30); st_stream(wk + (28) * 32, V31); st_stream(wk + (29) * 32, V32); st_stream(wk + (30) * 32, V33);
        st_stream(wk + (31) * 32, L.a02); st_stream(wk + (32) * 32, L.a12); st_stream(wk + (33) * 32, L.r0); st_stream(wk + (34) * 32, L.r1);
        st_stream(wk + (35) * 32, L.r2); st_stream(wk + (36) * 32, L.b00); st_stream(wk + (37) * 32, L.b10); st_stream(wk + (38) * 32, L.s0);
        st_stream(wk + (0) * 32, hk);
    }

    // ---- forward sweep (core/ddp.py:412-425) + gradient accumulation ---------
    R *dXo = a.dX + i * (4LL * (T + 1));
    R *dVo = a.dV + i * (2LL * T);
    R *dLo = a.dL + i * (4LL * (T + 1));
    const bool want_g = a.g_theta != nullptr;
    const R aw = alpha_grad_weight(th);
    GradAcc<R> G;
    R d0 = R(0), d1 = R(0), d2 = R(0), d3 = R(0);
    bool bad = false;
    // carried from step k-1 for the dynamics term of the accumulation
    R p_px = 0, p_py = 0, p_pt = 0, p_pb = 0, p_u0 = 0, p_u1 = 0, p_r0 = 0, p_r1 = 0, p_r2 = 0, p_q0 = 0, p_q1 = 0;
    R p_d0 = 0, p_d1 = 0, p_d2 = 0, p_d3 = 0, p_du0 = 0, p_du1 = 0, p_h = 0;
    // wk, Xk, Vk stand at step 0 after the backward sweep
    const R *xrk = xrp, *urk = urp, *Xbk = Xb;
    R *dXk = dXo, *dVk = dVo, *dLk = dLo, *gxk = gxo, *guk = guo;
    for (int k = 0; k < T; ++k) {
        const R K00 = ld_stream(wk + (1) * 32), K01 = ld_stream(wk + (2) * 32), K02 = ld_stream(wk + (3) * 32), K03 = ld_stream(wk + (4) * 32);
        const R K10 = ld_stream(wk + (5) * 32), K11 = ld_stream(wk + (6) * 32), K12 = ld_stream(wk + (7) * 32), K13 = ld_stream(wk + (8) * 32);
        const R k0 = ld_stream(wk + (9) * 32), k1 = ld_stream(wk + (10) * 32);
        R px, py, pt, pb, u0, u1;
        ld4(Xk, px, py, pt, pb);
        ld2(Vk, u0, u1);
        bool a0, a1;
        active_u(E, u0, u1, a0, a1);
        R du0 = k0 + m::fma_(K03, d3, m::fma_(K02, d2, m::fma_(K01, d1, K00 * d0)));
        R du1 = k1 + m::fma_(K13, d3, m::fma_(K12, d2, m::fma_(K11, d1, K10 * d0)));
        if (a0) du0 = R(0);
        if (a1) du1 = R(0);
        // delta_lambda_k = tilde V_x,k + V_xx,k delta_x_k
        const R l0 = ld_stream(wk + (11) * 32) + m::fma_(ld_stream(wk + (18) * 32), d3, m::fma_(ld_stream(wk + (17) * 32), d2, m::fma_(ld_stream(wk + (16) * 32), d1, ld_stream(wk + (15) * 32) * d0)));
        const R l1 = ld_stream(wk + (12) * 32) + m::fma_(ld_stream(wk + (22) * 32), d3, m::fma_(ld_stream(wk + (21) * 32), d2, m::fma_(ld_stream(wk + (20) * 32), d1, ld_stream(wk + (19) * 32) * d0)));
        const R l2 = ld_stream(wk + (13) * 32) + m::fma_(ld_stream(wk + (26) * 32), d3, m::fma_(ld_stream(wk + (25) * 32), d2, m::fma_(ld_stream(wk + (24) * 32), d1, ld_stream(wk + (23) * 32) * d0)));
        const R l3 = ld_stream(wk + (14) * 32) + m::fma_(ld_stream(wk + (30) * 32), d3, m::fma_(ld_stream(wk + (29) * 32), d2, m::fma_(ld_stream(wk + (28) * 32), d1, ld_stream(wk + (27) * 32) * d0)));
        st4(dXk, d0, d1, d2, d3);
        st2(dVk, du0, du1);
        st4(dLk, l0, l1, l2, l3);
        bad |= !finite_((l0 + l1) + (l2 + l3) + (du0 + du1));
        const R hk = ld_stream(wk + (0) * 32);
        R r0, r1, r2, q0, q1;
        if (xrp) {
            r0 = xrk[0]; r1 = xrk[1]; r2 = xrk[2];
            ld2(urk, q0, q1);
            xrk += xrow;
            urk += 2;
        } else {
            const R *tp = a.target + i * a.target_stride;
            r0 = tp[0]; r1 = tp[1]; r2 = tp[2];
            q0 = R(0); q1 = R(0);
        }
        if (want_g && k > 0)
            accumulate_stage(th, jc, G, aw, p_px, p_py, p_pt, p_pb, p_u0, p_u1, p_r0, p_r1, p_r2, p_q0, p_q1, p_d0, p_d1, p_d2,
                             p_d3, p_du0, p_du1, l3, p_h, hk);
        if (gxo) {
            gxk[0] = -(R(2) * th.Q0) * d0; gxk[1] = -(R(2) * th.Q1) * d1; gxk[2] = -(R(2) * th.Q2) * d2;
            gxk += 3;
        }
        if (guo) {
            st2(guk, -(R(2) * th.R0) * du0, -(R(2) * th.R1) * du1);
            guk += 2;
        }
        if (a.loss && Xb) {
            R b0, b1, b2, b3;
            ld4(Xbk, b0, b1, b2, b3);
            Xbk += 4;
            const R e0 = px - b0, e1 = py - b1, e2 = pt - b2;
            loss_t += m::fma_(e2, e2, m::fma_(e1, e1, e0 * e0));
            loss_b = m::fma_(pb, pb, loss_b);
        }
        p_px = px; p_py = py; p_pt = pt; p_pb = pb; p_u0 = u0; p_u1 = u1;
        p_r0 = r0; p_r1 = r1; p_r2 = r2; p_q0 = q0; p_q1 = q1;
        p_d0 = d0; p_d1 = d1; p_d2 = d2; p_d3 = d3; p_du0 = du0; p_du1 = du1; p_h = hk;
        // delta_x_{k+1} = A delta_x_k + B delta_u_k
        const R a02 = ld_stream(wk + (31) * 32), a12 = ld_stream(wk + (32) * 32), R0 = ld_stream(wk + (33) * 32), R1 = ld_stream(wk + (34) * 32), R2 = ld_stream(wk + (35) * 32);
        const R b00 = ld_stream(wk + (36) * 32), b10 = ld_stream(wk + (37) * 32), s0 = ld_stream(wk + (38) * 32);
        const R n0 = m::fma_(b00, du0, m::fma_(a02, d2, d0));
        const R n1 = m::fma_(b10, du0, m::fma_(a12, d2, d1));
        const R n2 = m::fma_(dt, du1, d2);
        const R n3 = m::fma_(s0, du0, m::fma_(gam, d3, m::fma_(R2, d2, m::fma_(R1, d1, R0 * d0))));
        d0 = n0; d1 = n1; d2 = n2; d3 = n3;
        wk += SENS_ROWS * 32;
        Xk += 4;
        Vk += 2;
        dXk += 4;
        dVk += 2;
        dLk += 4;
    }
#undef WK
    {
        // terminal: delta_lambda_T = g_x,T + phi_xx delta_x_T   (core/ddp.py:425; V_xx,T = diag(2Qf, 2qb))
        R px, py, pt, pb;
        ld4(X + 4 * T, px, py, pt, pb);
        R g0, g1, g2, g3;
        upper_gx(T, px, py, pt, pb, g0, g1, g2, g3);
        const R l0 = m::fma_(R(2) * th.Qf0, d0, g0), l1 = m::fma_(R(2) * th.Qf1, d1, g1);
        const R l2 = m::fma_(R(2) * th.Qf2, d2, g2), l3 = m::fma_(R(2) * th.qb, d3, g3);
        st4(dXo + 4 * T, d0, d1, d2, d3);
        st4(dLo + 4 * T, l0, l1, l2, l3);
        bad |= !finite_((l0 + l1) + (l2 + l3) + (d0 + d1) + (d2 + d3));
        R r0, r1, r2;
        if (xrp) {
            const R *xr = xrp + T * xrow;
            r0 = xr[0]; r1 = xr[1]; r2 = xr[2];
        } else {
            const R *tp = a.target + i * a.target_stride;
            r0 = tp[0]; r1 = tp[1]; r2 = tp[2];
        }
        if (want_g) {
            const R hT = WS(ROW(T, 0));
            accumulate_stage(th, jc, G, aw, p_px, p_py, p_pt, p_pb, p_u0, p_u1, p_r0, p_r1, p_r2, p_q0, p_q1, p_d0, p_d1, p_d2,
                             p_d3, p_du0, p_du1, l3, p_h, hT);
            G.gQf0 = (R(2) * (px - r0)) * d0;
            G.gQf1 = (R(2) * (py - r1)) * d1;
            G.gQf2 = (R(2) * (pt - r2)) * d2;
            G.gqb = m::fma_(R(2) * pb, d3, G.gqb);
            store_grads(a.g_theta + i * DTMPC_NTHETA, G);
        }
        if (gxo) {
            R *o = gxo + 3 * T;
            o[0] = -(R(2) * th.Qf0) * d0; o[1] = -(R(2) * th.Qf1) * d1; o[2] = -(R(2) * th.Qf2) * d2;
        }
        if (a.loss && Xb) {
            R b0, b1, b2, b3;
            ld4(Xb + 4 * T, b0, b1, b2, b3);
            const R e0 = px - b0, e1 = py - b1, e2 = pt - b2;
            loss_t += m::fma_(e2, e2, m::fma_(e1, e1, e0 * e0));
            loss_b = m::fma_(pb, pb, loss_b);
            a.loss[i] = loss_t + loss_b;
        }
    }
    if (a.status && bad) a.status[i] |= DTMPC_ST_NONFINITE << 8;
#undef WS
#undef ROW
}

// ift_gradient alone (reference core/ift.py:35-92) for caller-supplied delta_X, delta_V, delta_lambda.
template <typename R>
__global__ void __launch_bounds__(SENS_BLOCK) ift_accumulate_kernel(const __grid_constant__ SensParams<R> P) {
    const long long i = (long long)blockIdx.x * SENS_BLOCK + threadIdx.x;
    if (i >= P.B) return;
    const EnvT<R> &E = P.env;
    const SensArgs<R> &a = P.io;
    const int T = P.ocp.T;
    Theta<R> th;
    load_theta(th, a.theta + i * a.theta_stride, P.ocp);
    const R *X = a.X + i * (4LL * (T + 1));
    const R *V = a.V + i * (2LL * T);
    const R *dX = a.dX + i * (4LL * (T + 1));
    const R *dV = a.dV + i * (2LL * T);
    const R *dL = a.dL + i * (4LL * (T + 1));
    const R aw = alpha_grad_weight(th);
    const JacConsts<R> jc = jac_consts(th);
    GradAcc<R> G;
    R hc, gx, gy;
    eval_h<R, false>(E, X[0], X[1], hc, gx, gy);
    for (int k = 0; k <= T; ++k) {
        R r0, r1, r2, q0 = R(0), q1 = R(0);
        if (a.Xref) {
            const R *xr = a.Xref + (i * (T + 1) + k) * a.xref_row;
            r0 = xr[0]; r1 = xr[1]; r2 = xr[2];
            if (k < T) {
                q0 = a.Uref[i * (2LL * T) + 2 * k];
                q1 = a.Uref[i * (2LL * T) + 2 * k + 1];
            }
        } else {
            const R *tp = a.target + i * a.target_stride;
            r0 = tp[0]; r1 = tp[1]; r2 = tp[2];
        }
        const R px = X[4 * k], py = X[4 * k + 1], pt = X[4 * k + 2], pb = X[4 * k + 3];
        const R d0 = dX[4 * k], d1 = dX[4 * k + 1], d2 = dX[4 * k + 2], d3 = dX[4 * k + 3];
        if (k < T) {
            const R u0 = V[2 * k], u1 = V[2 * k + 1];
            const R du0 = dV[2 * k], du1 = dV[2 * k + 1];
            R hn;
            eval_h<R, false>(E, X[4 * (k + 1)], X[4 * (k + 1) + 1], hn, gx, gy);
            accumulate_stage(th, jc, G, aw, px, py, pt, pb, u0, u1, r0, r1, r2, q0, q1, d0, d1, d2, d3, du0, du1,
                             dL[4 * (k + 1) + 3], hc, hn);
            hc = hn;
            if (a.gXref) {
                R *o = a.gXref + (i * (T + 1) + k) * 3;
                o[0] = -(R(2) * th.Q0) * d0; o[1] = -(R(2) * th.Q1) * d1; o[2] = -(R(2) * th.Q2) * d2;
            }
            if (a.gUref) {
                R *o = a.gUref + i * (2LL * T) + 2 * k;
                o[0] = -(R(2) * th.R0) * du0; o[1] = -(R(2) * th.R1) * du1;
            }
        } else {
            G.gQf0 = (R(2) * (px - r0)) * d0;
            G.gQf1 = (R(2) * (py - r1)) * d1;
            G.gQf2 = (R(2) * (pt - r2)) * d2;
            G.gqb = M<R>::fma_(R(2) * pb, d3, G.gqb);
            if (a.gXref) {
                R *o = a.gXref + (i * (T + 1) + T) * 3;
                o[0] = -(R(2) * th.Qf0) * d0; o[1] = -(R(2) * th.Qf1) * d1; o[2] = -(R(2) * th.Qf2) * d2;
            }
        }
    }
    if (a.g_theta) store_grads(a.g_theta + i * DTMPC_NTHETA, G);
}

}  // namespace dtmpc
