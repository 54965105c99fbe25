// dtmpc_solve.cuh -- persistent batched iLQR kernel (reference core/ddp.py:102-307).
//
// Mapping: one problem per thread.  The sweep is serial in T and the blocks are
// tiny (nx = 4, nu = 2), so the Riccati blocks live in registers and all
// parallelism is across the batch.  Iteration counts differ per instance
// (reference: 2..20), so lanes are persistent workers: when a lane's OCP
// converges it writes its result and pulls the next instance from a global
// counter.  A warp therefore always executes the same phase (rollout /
// backward / line search / commit) with lanes at different iterations -- and,
// for the two-layer solve, in different OCPs (nominal vs auxiliary), which is
// the same code with different per-lane parameters.
//
// Working set per lane (trajectory X,V; gains K,k; tracked reference) lives in
// a slot-major structure-of-arrays scratch:  row r of lane s at ws[r*S + s],
// so every per-timestep access of a warp is one contiguous 256 B segment.
// Results leave the kernel instance-major ([B,T+1,4] ...): one state is one
// 32 B sector, which keeps scattered writes from refilled lanes sector-exact.
#pragma once
#include "dtmpc_math.cuh"

namespace dtmpc {

template <typename R> struct SolveArgs {
    const R *x0, *V_init, *theta, *target, *Xref, *Uref;
    long long theta_stride, target_stride;
    R *X, *V, *K, *kff;
    int *iters, *status, *alpha_idx;
    R *costs, *margins;
};

template <typename R> struct SolveParams {
    EnvT<R> env;
    OcpT<R> ocp[2];
    SolveArgs<R> io[2];
    long long B;
    int n_stage;  // 1: single OCP, 2: nominal then auxiliary tracking the nominal result
    unsigned long long *counter;
    R *ws;
    long long S;  // number of lanes (scratch stride)
};

constexpr int SOLVE_BLOCK = 128;

// scratch rows per lane
__host__ __device__ inline long long solve_ws_rows(int T) {
    return 4LL * (T + 1) + 2LL * T + 8LL * T + 2LL * T + 3LL * (T + 1) + 2LL * T;
}

template <typename R> struct Cand {
    State<R> x;
    R J;
};

template <typename R>
__global__ void __launch_bounds__(SOLVE_BLOCK) ilqr_kernel(const __grid_constant__ SolveParams<R> P) {
    using m = M<R>;
    const EnvT<R> &E = P.env;
    const long long S = P.S;
    const long long slot = (long long)blockIdx.x * SOLVE_BLOCK + threadIdx.x;
    const int T = P.ocp[0].T;
    R *const wX = P.ws + slot;
    R *const wV = wX + 4LL * (T + 1) * S;
    R *const wK = wV + 2LL * T * S;
    R *const wk = wK + 8LL * T * S;
    R *const wXr = wk + 2LL * T * S;
    R *const wUr = wXr + 3LL * (T + 1) * S;
    const unsigned FULL = 0xffffffffu;

    // ---- per-lane solver state -------------------------------------------
    long long inst = -1;  // -1: needs work, -2: queue exhausted
    int stage = 0, it = 0, status = 0;
    bool fresh = false, have_prev = false, use_ref = false;
    R prev = R(0), Jcur = R(0), B0 = R(0);
    R x00 = R(0), x01 = R(0), x02 = R(0), x03 = R(0);
    R tg0 = R(0), tg1 = R(0), tg2 = R(0);
    R tol = R(0), reg = R(0);
    int max_iter = 0, n_alpha = 0;
    Theta<R> th;

#define WS(base, row) (base)[(long long)(row) * S]
#define REF_AT(k, r0, r1, r2)                  \
    do {                                       \
        if (use_ref) {                         \
            r0 = WS(wXr, 3 * (k) + 0);         \
            r1 = WS(wXr, 3 * (k) + 1);         \
            r2 = WS(wXr, 3 * (k) + 2);         \
        } else {                               \
            r0 = tg0; r1 = tg1; r2 = tg2;      \
        }                                      \
    } while (0)
#define UREF_AT(k, q0, q1)                     \
    do {                                       \
        if (use_ref) {                         \
            q0 = WS(wUr, 2 * (k) + 0);         \
            q1 = WS(wUr, 2 * (k) + 1);         \
        } else {                               \
            q0 = R(0); q1 = R(0);              \
        }                                      \
    } while (0)

    // Loads everything a lane needs to start OCP `stage` of instance `inst`
    // (reference core/ddp.py:125-129: V = clamp(V_init)).
    auto setup_stage = [&]() {
        const OcpT<R> &o = P.ocp[stage];
        const SolveArgs<R> &a = P.io[stage];
        load_theta(th, a.theta + inst * a.theta_stride, o);
        tol = o.tol;
        reg = o.reg;
        max_iter = o.max_iter;
        n_alpha = o.n_alpha;
        const R *x0p = a.x0 + inst * 4;
        x00 = x0p[0]; x01 = x0p[1]; x02 = x0p[2]; x03 = x0p[3];
        B0 = barrier_of_state(E, th, x00, x01);
        const R *vp = a.V_init + inst * (2LL * T);
        for (int k = 0; k < T; ++k) {
            R u0 = vp[2 * k], u1 = vp[2 * k + 1];
            clamp_u(E, u0, u1);
            WS(wV, 2 * k) = u0;
            WS(wV, 2 * k + 1) = u1;
        }
        if (stage == 1 && P.n_stage == 2) {
            use_ref = true;  // reference already staged by finish_stage()
        } else if (a.Xref != nullptr) {
            use_ref = true;
            const R *xr = a.Xref + inst * (3LL * (T + 1));
            const R *ur = a.Uref + inst * (2LL * T);
            for (int k = 0; k <= T; ++k) {
                WS(wXr, 3 * k + 0) = xr[3 * k + 0];
                WS(wXr, 3 * k + 1) = xr[3 * k + 1];
                WS(wXr, 3 * k + 2) = xr[3 * k + 2];
            }
            for (int k = 0; k < T; ++k) {
                WS(wUr, 2 * k + 0) = ur[2 * k + 0];
                WS(wUr, 2 * k + 1) = ur[2 * k + 1];
            }
        } else {
            use_ref = false;
            const R *tp = a.target + inst * a.target_stride;
            tg0 = tp[0]; tg1 = tp[1]; tg2 = tp[2];
        }
        it = 0;
        have_prev = false;
        fresh = true;
    };

    // Writes the lane's result instance-major and either chains into the auxiliary OCP or frees the lane.
    auto finish_stage = [&](int st_bits) {
        const SolveArgs<R> &a = P.io[stage];
        R *Xo = a.X + inst * (4LL * (T + 1));
        R *Vo = a.V + inst * (2LL * T);
        for (int k = 0; k <= T; ++k) {
            Xo[4 * k + 0] = WS(wX, 4 * k + 0);
            Xo[4 * k + 1] = WS(wX, 4 * k + 1);
            Xo[4 * k + 2] = WS(wX, 4 * k + 2);
            Xo[4 * k + 3] = WS(wX, 4 * k + 3);
        }
        for (int k = 0; k < T; ++k) {
            Vo[2 * k + 0] = WS(wV, 2 * k + 0);
            Vo[2 * k + 1] = WS(wV, 2 * k + 1);
        }
        if (a.K != nullptr) {
            R *Ko = a.K + inst * (8LL * T);
            for (int r = 0; r < 8 * T; ++r) Ko[r] = WS(wK, r);
        }
        if (a.kff != nullptr) {
            R *ko = a.kff + inst * (2LL * T);
            for (int r = 0; r < 2 * T; ++r) ko[r] = WS(wk, r);
        }
        if (a.iters != nullptr) a.iters[inst] = it;
        status |= st_bits << (4 * stage);
        if (P.n_stage == 2 && stage == 0 && st_bits == 0) {
            // the nominal solution becomes the tracked reference (core/tube_mpc.py:864-865)
            for (int k = 0; k <= T; ++k) {
                WS(wXr, 3 * k + 0) = WS(wX, 4 * k + 0);
                WS(wXr, 3 * k + 1) = WS(wX, 4 * k + 1);
                WS(wXr, 3 * k + 2) = WS(wX, 4 * k + 2);
            }
            for (int r = 0; r < 2 * T; ++r) WS(wUr, r) = WS(wV, r);
            stage = 1;
            setup_stage();
        } else {
            P.io[0].status[inst] = status;
            inst = -1;
        }
    };

    for (;;) {
        // ---- work acquisition --------------------------------------------
        if (inst == -1) {
            unsigned long long t = atomicAdd(P.counter, 1ULL);
            if (t >= (unsigned long long)P.B) {
                inst = -2;
            } else {
                inst = (long long)t;
                stage = 0;
                status = 0;
                setup_stage();
            }
        }
        if (__all_sync(FULL, inst == -2)) break;
        const bool live = inst >= 0;

        // ---- phase R: initial rollout of freshly loaded OCPs (core/ddp.py:131) ----
        if (__any_sync(FULL, live && fresh)) {
            if (live && fresh) {
                State<R> x{x00, x01, x02, x03, B0};
                R J = R(0);
                WS(wX, 0) = x00; WS(wX, 1) = x01; WS(wX, 2) = x02; WS(wX, 3) = x03;
                for (int k = 0; k < T; ++k) {
                    R u0 = WS(wV, 2 * k), u1 = WS(wV, 2 * k + 1);
                    R r0, r1, r2, q0, q1;
                    REF_AT(k, r0, r1, r2);
                    UREF_AT(k, q0, q1);
                    J += stage_cost(th, x.px, x.py, x.th, x.b, u0, u1, r0, r1, r2, q0, q1);
                    step(E, th, x, u0, u1);
                    WS(wX, 4 * (k + 1) + 0) = x.px;
                    WS(wX, 4 * (k + 1) + 1) = x.py;
                    WS(wX, 4 * (k + 1) + 2) = x.th;
                    WS(wX, 4 * (k + 1) + 3) = x.b;
                }
                R r0, r1, r2;
                REF_AT(T, r0, r1, r2);
                J += terminal_cost(th, x.px, x.py, x.th, x.b, r0, r1, r2);
                Jcur = J;
                fresh = false;
            }
        }

        bool bad = false;
        if (live) {
            // ---- phase B: linearise + backward Riccati sweep (core/ddp.py:172-254) ----
            R Vx0, Vx1, Vx2, Vx3;
            R V00, V01, V02, V03, V10, V11, V12, V13, V20, V21, V22, V23, V30, V31, V32, V33;
            R gnx, gny;
            {
                R px = WS(wX, 4 * T + 0), py = WS(wX, 4 * T + 1), pt = WS(wX, 4 * T + 2), pb = WS(wX, 4 * T + 3);
                R r0, r1, r2;
                REF_AT(T, r0, r1, r2);
                // terminal derivatives: core/cost_derivs.py:132-146 + barrier entries core/tube_mpc.py:839-841
                Vx0 = (R(2) * th.Qf0) * (px - r0);
                Vx1 = (R(2) * th.Qf1) * (py - r1);
                Vx2 = (R(2) * th.Qf2) * heading_err_for_derivs(th, pt, r2);
                Vx3 = (R(2) * th.qb) * pb;
                V00 = R(2) * th.Qf0; V11 = R(2) * th.Qf1; V22 = R(2) * th.Qf2; V33 = R(2) * th.qb;
                V01 = V02 = V03 = V10 = V12 = V13 = V20 = V21 = V23 = V30 = V31 = V32 = R(0);
                R hT;
                barrier_grad(E, th, px, py, gnx, gny, hT);
            }
            const R lxx0 = R(2) * th.Q0, lxx1 = R(2) * th.Q1, lxx2 = R(2) * th.Q2, lxx3 = R(2) * th.qb;
            const R luu0 = R(2) * th.R0, luu1 = R(2) * th.R1;
            const R gam = th.gamma, dt = E.dt;
            for (int k = T - 1; k >= 0; --k) {
                const R px = WS(wX, 4 * k + 0), py = WS(wX, 4 * k + 1), pt = WS(wX, 4 * k + 2), pb = WS(wX, 4 * k + 3);
                const R u0 = WS(wV, 2 * k), u1 = WS(wV, 2 * k + 1);
                R r0, r1, r2, q0, q1;
                REF_AT(k, r0, r1, r2);
                UREF_AT(k, q0, q1);
                R gcx, gcy, hk;
                barrier_grad(E, th, px, py, gcx, gcy, hk);
                Lin<R> L;
                linearize(E, th, pt, u0, gnx, gny, gcx, gcy, L);
                gnx = gcx;
                gny = gcy;
                // stage derivatives, core/cost_derivs.py:58-76, 110-129
                const R lx0 = lxx0 * (px - r0), lx1 = lxx1 * (py - r1);
                const R lx2 = lxx2 * heading_err_for_derivs(th, pt, r2), lx3 = lxx3 * pb;
                const R lu0 = luu0 * (u0 - q0), lu1 = luu1 * (u1 - q1);

                // M = V_xx A   (A = [1 0 a02 0; 0 1 a12 0; 0 0 1 0; r0 r1 r2 gamma])
#define ROW_M(i)                                                                              \
    const R M##i##0 = m::fma_(V##i##3, L.r0, V##i##0);                                        \
    const R M##i##1 = m::fma_(V##i##3, L.r1, V##i##1);                                        \
    const R M##i##2 = m::fma_(V##i##3, L.r2, m::fma_(V##i##1, L.a12, m::fma_(V##i##0, L.a02, V##i##2))); \
    const R M##i##3 = V##i##3 * gam;
                ROW_M(0) ROW_M(1) ROW_M(2) ROW_M(3)
#undef ROW_M
                // Q_xx = l_xx + A^T M
#define COL_Q(j)                                                                              \
    R Qxx0##j = m::fma_(L.r0, M3##j, M0##j);                                                  \
    R Qxx1##j = m::fma_(L.r1, M3##j, M1##j);                                                  \
    R Qxx2##j = m::fma_(L.r2, M3##j, m::fma_(L.a12, M1##j, m::fma_(L.a02, M0##j, M2##j)));    \
    R Qxx3##j = gam * M3##j;
                COL_Q(0) COL_Q(1) COL_Q(2) COL_Q(3)
#undef COL_Q
                Qxx00 += lxx0; Qxx11 += lxx1; Qxx22 += lxx2; Qxx33 += lxx3;
                // Q_ux = B^T M   (B = [b00 0; b10 0; 0 dt; s0 0])
#define COL_U(j)                                                                              \
    const R Qux0##j = m::fma_(L.s0, M3##j, m::fma_(L.b10, M1##j, L.b00 * M0##j));             \
    const R Qux1##j = dt * M2##j;
                COL_U(0) COL_U(1) COL_U(2) COL_U(3)
#undef COL_U
                // V_xx B columns, Q_uu = l_uu + B^T V_xx B
                const R w0 = m::fma_(V03, L.s0, m::fma_(V01, L.b10, V00 * L.b00));
                const R w1 = m::fma_(V13, L.s0, m::fma_(V11, L.b10, V10 * L.b00));
                const R w2 = m::fma_(V23, L.s0, m::fma_(V21, L.b10, V20 * L.b00));
                const R w3 = m::fma_(V33, L.s0, m::fma_(V31, L.b10, V30 * L.b00));
                const R y0 = V02 * dt, y1 = V12 * dt, y2 = V22 * dt, y3 = V32 * dt;
                const R Quu00 = luu0 + m::fma_(L.s0, w3, m::fma_(L.b10, w1, L.b00 * w0));
                const R Quu01 = m::fma_(L.s0, y3, m::fma_(L.b10, y1, L.b00 * y0));
                const R Quu10 = dt * w2;
                const R Quu11 = luu1 + dt * y2;
                // Q_x = l_x + A^T V_x, Q_u = l_u + B^T V_x
                const R Qx0 = lx0 + m::fma_(L.r0, Vx3, Vx0);
                const R Qx1 = lx1 + m::fma_(L.r1, Vx3, Vx1);
                const R Qx2 = lx2 + m::fma_(L.r2, Vx3, m::fma_(L.a12, Vx1, m::fma_(L.a02, Vx0, Vx2)));
                const R Qx3 = lx3 + gam * Vx3;
                const R Qu0 = lu0 + m::fma_(L.s0, Vx3, m::fma_(L.b10, Vx1, L.b00 * Vx0));
                const R Qu1 = lu1 + dt * Vx2;
                // gains: K = -(Q_uu + reg I)^-1 Q_ux, k = -(Q_uu + reg I)^-1 Q_u   (core/ddp.py:239-244)
                LU2<R> F;
                lu2_factor(F, Quu00 + reg, Quu01, Quu10, Quu11 + reg);
                R K00, K01, K02, K03, K10, K11, K12, K13, k0, k1;
                lu2_solve(F, Qux00, Qux10, K00, K10);
                lu2_solve(F, Qux01, Qux11, K01, K11);
                lu2_solve(F, Qux02, Qux12, K02, K12);
                lu2_solve(F, Qux03, Qux13, K03, K13);
                lu2_solve(F, Qu0, Qu1, k0, k1);
                K00 = -K00; K01 = -K01; K02 = -K02; K03 = -K03;
                K10 = -K10; K11 = -K11; K12 = -K12; K13 = -K13;
                k0 = -k0; k1 = -k1;
                WS(wK, 8 * k + 0) = K00; WS(wK, 8 * k + 1) = K01; WS(wK, 8 * k + 2) = K02; WS(wK, 8 * k + 3) = K03;
                WS(wK, 8 * k + 4) = K10; WS(wK, 8 * k + 5) = K11; WS(wK, 8 * k + 6) = K12; WS(wK, 8 * k + 7) = K13;
                WS(wk, 2 * k + 0) = k0; WS(wk, 2 * k + 1) = k1;
                // V_x = Q_x + K^T Q_uu k + K^T Q_u + Q_xu k   (un-regularised Q_uu, core/ddp.py:251)
                const R t0 = m::fma_(Quu01, k1, Quu00 * k0) + Qu0;
                const R t1 = m::fma_(Quu11, k1, Quu10 * k0) + Qu1;
                Vx0 = Qx0 + m::fma_(K10, t1, K00 * t0) + m::fma_(Qux10, k1, Qux00 * k0);
                Vx1 = Qx1 + m::fma_(K11, t1, K01 * t0) + m::fma_(Qux11, k1, Qux01 * k0);
                Vx2 = Qx2 + m::fma_(K12, t1, K02 * t0) + m::fma_(Qux12, k1, Qux02 * k0);
                Vx3 = Qx3 + m::fma_(K13, t1, K03 * t0) + m::fma_(Qux13, k1, Qux03 * k0);
                // V_xx = Q_xx + K^T Q_uu K + K^T Q_ux + Q_xu K   (never symmetrised, core/ddp.py:252)
#define COL_W(j)                                                              \
    const R W0##j = m::fma_(Quu01, K1##j, Quu00 * K0##j) + Qux0##j;           \
    const R W1##j = m::fma_(Quu11, K1##j, Quu10 * K0##j) + Qux1##j;
                COL_W(0) COL_W(1) COL_W(2) COL_W(3)
#undef COL_W
#define ENT_V(i, j) V##i##j = Qxx##i##j + m::fma_(K1##i, W1##j, K0##i * W0##j) + m::fma_(Qux1##i, K1##j, Qux0##i * K0##j);
                ENT_V(0, 0) ENT_V(0, 1) ENT_V(0, 2) ENT_V(0, 3)
                ENT_V(1, 0) ENT_V(1, 1) ENT_V(1, 2) ENT_V(1, 3)
                ENT_V(2, 0) ENT_V(2, 1) ENT_V(2, 2) ENT_V(2, 3)
                ENT_V(3, 0) ENT_V(3, 1) ENT_V(3, 2) ENT_V(3, 3)
#undef ENT_V
                // non-finite guard (reference raises: core/ddp.py:225-254)
                R chk = (Vx0 + Vx1) + (Vx2 + Vx3) + (k0 + k1) + (K00 + K01) + (K02 + K03) + (K10 + K11) + (K12 + K13) +
                        (V00 + V11) + (V22 + V33) + (V01 + V02) + (V03 + V12) + (V13 + V23) + (V10 + V20) + (V30 + V21) +
                        (V31 + V32);
                bad |= !finite_(chk);
            }

            // ---- phase F: forward pass, every alpha evaluated (core/ddp.py:256-296) ----
            Cand<R> cs[DTMPC_MAX_ALPHAS];
            const OcpT<R> &o = P.ocp[stage];
            for (int c = 0; c < n_alpha; ++c) {
                cs[c].x = State<R>{x00, x01, x02, x03, B0};
                cs[c].J = R(0);
            }
            for (int k = 0; k < T; ++k) {
                const R X0 = WS(wX, 4 * k + 0), X1 = WS(wX, 4 * k + 1), X2 = WS(wX, 4 * k + 2), X3 = WS(wX, 4 * k + 3);
                const R v0 = WS(wV, 2 * k), v1 = WS(wV, 2 * k + 1);
                const R K00 = WS(wK, 8 * k + 0), K01 = WS(wK, 8 * k + 1), K02 = WS(wK, 8 * k + 2), K03 = WS(wK, 8 * k + 3);
                const R K10 = WS(wK, 8 * k + 4), K11 = WS(wK, 8 * k + 5), K12 = WS(wK, 8 * k + 6), K13 = WS(wK, 8 * k + 7);
                const R k0 = WS(wk, 2 * k), k1 = WS(wk, 2 * k + 1);
                R r0, r1, r2, q0, q1;
                REF_AT(k, r0, r1, r2);
                UREF_AT(k, q0, q1);
                for (int c = 0; c < n_alpha; ++c) {
                    const R al = o.alphas[c];
                    if (al == R(0)) continue;  // current trajectory, cost already known (Jcur)
                    State<R> x = cs[c].x;
                    const R d0 = x.px - X0, d1 = x.py - X1, d2 = x.th - X2, d3 = x.b - X3;
                    const R du0 = k0 + m::fma_(K03, d3, m::fma_(K02, d2, m::fma_(K01, d1, K00 * d0)));
                    const R du1 = k1 + m::fma_(K13, d3, m::fma_(K12, d2, m::fma_(K11, d1, K10 * d0)));
                    R u0 = m::fma_(al, du0, v0), u1 = m::fma_(al, du1, v1);
                    clamp_u(E, u0, u1);
                    cs[c].J += stage_cost(th, x.px, x.py, x.th, x.b, u0, u1, r0, r1, r2, q0, q1);
                    step(E, th, x, u0, u1);
                    cs[c].x = x;
                }
            }
            int best_c = -1;
            R best = R(0), second = m::inf_();
            {
                R r0, r1, r2;
                REF_AT(T, r0, r1, r2);
                for (int c = 0; c < n_alpha; ++c) {
                    R Jc;
                    if (o.alphas[c] == R(0)) {
                        Jc = Jcur;
                    } else {
                        const State<R> &x = cs[c].x;
                        Jc = cs[c].J + terminal_cost(th, x.px, x.py, x.th, x.b, r0, r1, r2);
                    }
                    bad |= !finite_(Jc);
                    if (best_c < 0 || Jc < best) {  // strict '<', first wins (core/ddp.py:293)
                        if (best_c >= 0) second = m::min_(second, best);
                        best = Jc;
                        best_c = c;
                    } else {
                        second = m::min_(second, Jc);
                    }
                }
            }

            // ---- phase W: commit the winner (recomputed; candidates are not stored) ----
            const R aw = o.alphas[best_c];
            if (aw != R(0) && !bad) {
                State<R> x{x00, x01, x02, x03, B0};
                R J = R(0);
                R X0 = x00, X1 = x01, X2 = x02, X3 = x03;
                for (int k = 0; k < T; ++k) {
                    const R v0 = WS(wV, 2 * k), v1 = WS(wV, 2 * k + 1);
                    const R K00 = WS(wK, 8 * k + 0), K01 = WS(wK, 8 * k + 1), K02 = WS(wK, 8 * k + 2), K03 = WS(wK, 8 * k + 3);
                    const R K10 = WS(wK, 8 * k + 4), K11 = WS(wK, 8 * k + 5), K12 = WS(wK, 8 * k + 6), K13 = WS(wK, 8 * k + 7);
                    const R k0 = WS(wk, 2 * k), k1 = WS(wk, 2 * k + 1);
                    const R N0 = WS(wX, 4 * (k + 1) + 0), N1 = WS(wX, 4 * (k + 1) + 1), N2 = WS(wX, 4 * (k + 1) + 2),
                            N3 = WS(wX, 4 * (k + 1) + 3);
                    R r0, r1, r2, q0, q1;
                    REF_AT(k, r0, r1, r2);
                    UREF_AT(k, q0, q1);
                    const R d0 = x.px - X0, d1 = x.py - X1, d2 = x.th - X2, d3 = x.b - X3;
                    const R du0 = k0 + m::fma_(K03, d3, m::fma_(K02, d2, m::fma_(K01, d1, K00 * d0)));
                    const R du1 = k1 + m::fma_(K13, d3, m::fma_(K12, d2, m::fma_(K11, d1, K10 * d0)));
                    R u0 = m::fma_(aw, du0, v0), u1 = m::fma_(aw, du1, v1);
                    clamp_u(E, u0, u1);
                    WS(wV, 2 * k) = u0;
                    WS(wV, 2 * k + 1) = u1;
                    J += stage_cost(th, x.px, x.py, x.th, x.b, u0, u1, r0, r1, r2, q0, q1);
                    step(E, th, x, u0, u1);
                    WS(wX, 4 * (k + 1) + 0) = x.px;
                    WS(wX, 4 * (k + 1) + 1) = x.py;
                    WS(wX, 4 * (k + 1) + 2) = x.th;
                    WS(wX, 4 * (k + 1) + 3) = x.b;
                    X0 = N0; X1 = N1; X2 = N2; X3 = N3;
                }
                R r0, r1, r2;
                REF_AT(T, r0, r1, r2);
                J += terminal_cost(th, x.px, x.py, x.th, x.b, r0, r1, r2);
                Jcur = J;
            }

            // ---- bookkeeping: trace, tolerance break (core/ddp.py:301-305) ----
            const SolveArgs<R> &a = P.io[stage];
            if (!bad) {
                if (a.alpha_idx != nullptr) a.alpha_idx[inst * o.max_iter + it] = best_c;
                if (a.costs != nullptr) a.costs[inst * o.max_iter + it] = best;
                if (a.margins != nullptr) a.margins[inst * o.max_iter + it] = second - best;
                ++it;
            }
            const bool conv = have_prev && (m::abs_(prev - best) < tol);
            prev = best;
            have_prev = true;
            if (bad || conv || it >= max_iter) finish_stage(bad ? DTMPC_ST_NONFINITE : 0);
        }
    }
#undef WS
#undef REF_AT
#undef UREF_AT
}

}  // namespace dtmpc
