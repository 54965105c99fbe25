// dtmpc_solve.cuh -- batched iLQR (reference core/ddp.py:102-307) as phase-synchronous kernels.
//
// The sweep is serial in T and the blocks are tiny (nx = 4, nu = 2): all parallelism is across the batch.  Every iLQR
// iteration of the whole batch is two launches over the SLOTS that still hold a running instance:
//
//   ric_kernel     thread per slot                 linearisation (from the stored barrier gradients g_k) + backward
//                                                  Riccati sweep, register-resident V_x / V_xx, gains K, k
//   fwdc_kernel    thread per (slot, candidate)    forward line search, selection and commit.  Phase 1: every alpha != 0
//                                                  rolls out in its own lane, one warp per candidate, 32 slots per CTA,
//                                                  so the gains and the current trajectory are fetched once per CTA (L1)
//                                                  for all candidates; the candidate costs meet in shared memory.  Phase
//                                                  2: warp 0 takes the decision (argmin with strict <, tolerance break,
//                                                  trace) and reserves the survivors' slots of the next iteration.
//                                                  Phase 3: the warps roll the winner out once more (bit-identical to
//                                                  its candidate, dtmpc_math.cuh), one horizon segment each from the
//                                                  state the candidate checkpointed at the segment boundary, and write
//                                                  it either into the slot the instance holds in the NEXT iteration or,
//                                                  if the instance stops here, straight into the caller's result tensors
//
// Candidate trajectories are never stored: storing all of them made the line search HBM-bound (2.6 GB per iteration
// at B = 65,536, profiles/r01b), recomputing the winner costs 1/n_alpha more arithmetic and keeps one trajectory per
// instance.  The alpha = 0 candidate is the current trajectory bit for bit and costs nothing.
//
// Slots, not instances, index the big scratch arrays (trajectory, barrier gradients, gains), and the arrays are
// tile-major: 32 consecutive slots form a tile whose rows are contiguous, element (step k, row r) of slot s at
//     base[((s/32 * steps + k) * rows + r) * 32 + s%32].
// A warp reads one 256 B segment per row, every address of a sweep is one running pointer plus an immediate (a
// batch-sized row stride cost ~4 integer instructions per load in 64-bit address arithmetic), and the inputs of a step
// are one contiguous block.  (Interleaving the rows in PAIRS, so that a lane moves two values per 16-byte access, was
// measured in round 2: 6 % fewer instructions in the line search, 5 % fewer in the backward sweep -- and a SLOWER step,
// 14.7 vs 14.1 ms: the 512 B warp accesses hit L1 less often and the aligned register pairs cost spills.)
// The commit phase rewrites the whole trajectory every iteration
// anyway, so it writes it to the instance's next slot (double-buffered): the running instances always occupy slots
// 0 .. n_act-1.  Every access of a warp is then one full contiguous segment however few instances are left, where an
// index list over instance-indexed arrays cost up to 2.7x the traffic once a third of the batch had converged
// (profiles/r01c_iterations.txt).  Small per-instance scalars and the tracked reference stay instance-indexed (the
// reference in the same tile-major form).  Results leave instance-major ([B,T+1,4] ...: one state = one 32 B sector).
#pragma once
#include "dtmpc_math.cuh"

namespace dtmpc {

template <typename R> struct SolveArgs {
    const R *x0, *V_init, *theta, *target, *Xref, *Uref;
    long long theta_stride, target_stride;
    R *X, *V, *K, *kff;
    int *iters, *status, *alpha_idx;
    R *costs, *margins;
    int *fail_at;
};

template <typename R> struct SolveWs {
    int *counts;     // [2] number of occupied slots, per parity
    int *list[2];    // [B] instance held by each slot (ascending), per parity
    int *it;         // [B] completed iterations                         (per instance)
    int *flags;      // [B] bit 0: non-finite in the backward pass       (per instance)
    R *Jcur, *prev, *B0;  // [B] cost of the current trajectory, previous accepted cost, B(h(x0) - tight) (per instance)
    R *traj[2];      // [tiles][T+1][8][32] per slot and parity: x_k (4), u_k (2) and g_k = B'(h(x_k)) grad h(x_k), which is what
                     // the barrier row of the jacobian needs.  g is written by whoever rolls the trajectory out (init,
                     // commit); a separate linearisation kernel and its 8 stored jacobian entries per step made the
                     // backward pass HBM-bound (profiles/r01b)
    R *gains;        // [tiles][T][10][32] per slot: K_k (8) then k_k (2)
    R *ref;          // [tiles][T+1][5][32] tracked reference x (3), u (2) of the auxiliary OCP  (per instance)
    R *ckpt;         // [tiles][n_eval][n_seg-1][5][32] candidate states at the segment boundaries (DTMPC_CKPT_GLOBAL only)
};

template <typename R> struct SolveParams {
    EnvT<R> env;
    OcpT<R> ocp;
    SolveArgs<R> io;
    SolveWs<R> ws;
    long long B;
    int n_eval;                      // candidates that are actually rolled out (alpha != 0)
    int eval_idx[DTMPC_MAX_ALPHAS];  // their positions in the alpha list
    int use_ref;                     // 1: track ws.ref, 0: track io.target
    int ref_from_io;                 // 1: init copies io.Xref / io.Uref into ws.ref
    int write_ref;                   // 1: a stopping instance also writes ws.ref (nominal solution -> auxiliary reference)
    int status_shift;                // 0 (single / nominal) or 4 (auxiliary)
    int parity;                      // slot assignment / trajectory buffer read by this launch
    int eval_pos[DTMPC_MAX_ALPHAS];  // inverse of eval_idx (-1 for alpha = 0)
    int n_seg, seg_len;              // the commit sweep cuts the horizon into n_seg segments of seg_len steps
};

constexpr int LS_BLOCK = 128;

// rows per step of the tile-major arrays, and the element offset of (slot or instance) s in an array of `rows` rows
constexpr int TR = 8;    // trajectory: x 0-3, u 4-5, g 6-7
constexpr int GR = 10;   // gains: K 0-7 (row-major 2x4), k 8-9
constexpr int RR = 5;    // reference: x 0-2, u 3-4
constexpr int CR = 5;    // checkpoint: px, py, th, b, Bc
__host__ __device__ inline long long traj_rows(int T) { return (long long)TR * (T + 1); }
__host__ __device__ inline long long gain_rows(int T) { return (long long)GR * T; }
__host__ __device__ inline long long ref_rows(int T) { return (long long)RR * (T + 1); }
__host__ __device__ inline long long tile_off(long long s, long long rows) { return (s >> 5) * rows * 32 + (s & 31); }
#define EL(k, rows, r) ((((k) * (rows)) + (r)) * 32)  // element (step k, row r) relative to the slot's tile pointer

template <typename R> struct RefPt { R x0, x1, x2, u0, u1; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// UR: -1 = P.use_ref decides at run time, 0 / 1 = fixed at compile time (the line search is instantiated for both)
template <typename R, int UR = -1>
__device__ __forceinline__ void ref_at(const SolveParams<R> &P, const R *rk, bool with_u, R t0, R t1, R t2, RefPt<R> &r) {
    if (UR < 0 ? P.use_ref != 0 : UR != 0) {  // rk = ws.ref + tile_off(i, ref_rows(T)) + EL(k, RR, 0): the reference of step k
        r.x0 = rk[0 * 32];
        r.x1 = rk[1 * 32];
        r.x2 = rk[2 * 32];
        if (with_u) {
            r.u0 = rk[3 * 32];
            r.u1 = rk[4 * 32];
        } else {
            r.u0 = R(0); r.u1 = R(0);
        }
    } else {
        r.x0 = t0; r.x1 = t1; r.x2 = t2; r.u0 = R(0); r.u1 = R(0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// init: V = clamp(V_init), X = rollout(x0, V)  (core/ddp.py:125-131); every instance starts in slot i
// ---------------------------------------------------------------------------------------------------------------
template <typename R, int NOBS = -1>
__global__ void __launch_bounds__(LS_BLOCK) init_kernel(const __grid_constant__ SolveParams<R> P) {
    const long long i = (long long)blockIdx.x * LS_BLOCK + threadIdx.x;
    if (i == 0) {
        P.ws.counts[0] = (int)P.B;
        P.ws.counts[1] = 0;
    }
    if (i >= P.B) return;
    const EnvT<R> &E = P.env;
    const int T = P.ocp.T;
    Theta<R> th;
    load_theta(th, P.io.theta + i * P.io.theta_stride, P.ocp);
    R *ref_i = P.ws.ref + tile_off(i, ref_rows(T));
    if (P.ref_from_io) {
        const R *xr = P.io.Xref + i * (3LL * (T + 1));
        const R *ur = P.io.Uref + i * (2LL * T);
        for (int k = 0; k <= T; ++k)
            for (int c = 0; c < 3; ++c) ref_i[EL(k, RR, c)] = xr[3 * k + c];
        for (int k = 0; k < T; ++k)
            for (int c = 0; c < 2; ++c) ref_i[EL(k, RR, 3 + c)] = ur[2 * k + c];
    }
    R t0 = R(0), t1 = R(0), t2 = R(0);
    if (!P.use_ref) {
        const R *tp = P.io.target + i * P.io.target_stride;
        t0 = tp[0]; t1 = tp[1]; t2 = tp[2];
    }
    const R *x0p = P.io.x0 + i * 4;
    State<R> x{x0p[0], x0p[1], x0p[2], x0p[3], R(0)};
    x.Bc = barrier_of_state(E, th, x.px, x.py);
    P.ws.B0[i] = x.Bc;
    R *X = P.ws.traj[0] + tile_off(i, traj_rows(T));  // slot i of parity 0
    X[EL(0, TR, 0)] = x.px; X[EL(0, TR, 1)] = x.py; X[EL(0, TR, 2)] = x.th; X[EL(0, TR, 3)] = x.b;
    {
        R g0x, g0y, h0;
        barrier_grad<R, NOBS>(E, th, x.px, x.py, g0x, g0y, h0);
        X[EL(0, TR, 6)] = g0x; X[EL(0, TR, 7)] = g0y;
    }
    const R *vp = P.io.V_init + i * (2LL * T);
    R J = R(0);
    for (int k = 0; k < T; ++k) {
        R u0 = vp[2 * k], u1 = vp[2 * k + 1];
        clamp_u(E, u0, u1);
        X[EL(k, TR, 4)] = u0;
        X[EL(k, TR, 5)] = u1;
        RefPt<R> r;
        ref_at(P, ref_i + EL(k, RR, 0), true, t0, t1, t2, r);
        J += stage_cost(th, x.px, x.py, x.th, x.b, u0, u1, r.x0, r.x1, r.x2, r.u0, r.u1);
        R gx, gy;
        step_g<R, NOBS>(E, th, x, u0, u1, gx, gy);
        X[EL(k + 1, TR, 0)] = x.px;
        X[EL(k + 1, TR, 1)] = x.py;
        X[EL(k + 1, TR, 2)] = x.th;
        X[EL(k + 1, TR, 3)] = x.b;
        X[EL(k + 1, TR, 6)] = gx;
        X[EL(k + 1, TR, 7)] = gy;
    }
    RefPt<R> r;
    ref_at(P, ref_i + EL(T, RR, 0), false, t0, t1, t2, r);
    J += terminal_cost(th, x.px, x.py, x.th, x.b, r.x0, r.x1, r.x2);
    P.ws.Jcur[i] = J;
    P.ws.prev[i] = R(0);
    P.ws.it[i] = 0;
    P.ws.flags[i] = 0;
    P.ws.list[0][i] = (int)i;
    if (P.status_shift == 0) P.io.status[i] = 0;
    if (P.io.fail_at) {
        P.io.fail_at[2 * i] = -1;
        P.io.fail_at[2 * i + 1] = -1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// linearisation (core/ddp.py:172-199 -> core/systems/dubins_aug_jac.py:61-139) + backward Riccati sweep
// (core/ddp.py:208-254), thread per instance
// ---------------------------------------------------------------------------------------------------------------
// Both are HBM-bound streaming sweeps: occupancy beats a register double buffer.  Same-box A/B at B = 65,536, full
// mode: 255 regs + read-ahead 26.0 ms, 128 regs + no read-ahead 25.0 ms per tube solve.
#ifndef DTMPC_RIC_MAXNREG
#define DTMPC_RIC_MAXNREG 128
#endif
#ifndef DTMPC_RIC_MAXNREG_F32
#define DTMPC_RIC_MAXNREG_F32 128
#endif
#ifndef DTMPC_RIC_BLOCK
#define DTMPC_RIC_BLOCK 64
#endif
constexpr int RIC_BLOCK = DTMPC_RIC_BLOCK;

template <typename R>
__global__ void __maxnreg__(sizeof(R) == 8 ? DTMPC_RIC_MAXNREG : DTMPC_RIC_MAXNREG_F32) ric_kernel(const __grid_constant__ SolveParams<R> P) {
    using m = M<R>;
    const int n_act = P.ws.counts[P.parity];
    const int j = blockIdx.x * RIC_BLOCK + threadIdx.x;
    if (j == 0) P.ws.counts[P.parity ^ 1] = 0;   // fwdc_kernel of this iteration counts the survivors into it
    if (j >= n_act) return;
    const long long i = P.ws.list[P.parity][j];
    const EnvT<R> &E = P.env;
    const int T = P.ocp.T;
    Theta<R> th;
    load_theta(th, P.io.theta + i * P.io.theta_stride, P.ocp);
    R t0 = R(0), t1 = R(0), t2 = R(0);
    if (!P.use_ref) {
        const R *tp = P.io.target + i * P.io.target_stride;
        t0 = tp[0]; t1 = tp[1]; t2 = tp[2];
    }
    const R *X = P.ws.traj[P.parity] + tile_off(j, traj_rows(T));
    R *GK = P.ws.gains + tile_off(j, gain_rows(T));
    const R *ref_i = P.ws.ref + tile_off(i, ref_rows(T));
    const R reg = P.ocp.reg;

    R Vx0, Vx1, Vx2, Vx3;
    R V00, V01, V02, V03, V10, V11, V12, V13, V20, V21, V22, V23, V30, V31, V32, V33;
    {
        const R px = X[EL(T, TR, 0)], py = X[EL(T, TR, 1)], pt = X[EL(T, TR, 2)], pb = X[EL(T, TR, 3)];
        RefPt<R> r;
        ref_at(P, ref_i + EL(T, RR, 0), false, t0, t1, t2, r);
        // terminal derivatives: core/cost_derivs.py:132-146 + barrier entries core/tube_mpc.py:839-841
        Vx0 = (R(2) * th.Qf0) * (px - r.x0);
        Vx1 = (R(2) * th.Qf1) * (py - r.x1);
        Vx2 = (R(2) * th.Qf2) * heading_err_for_derivs(th, pt, r.x2);
        Vx3 = (R(2) * th.qb) * pb;
        V00 = R(2) * th.Qf0; V11 = R(2) * th.Qf1; V22 = R(2) * th.Qf2; V33 = R(2) * th.qb;
        V01 = V02 = V03 = V10 = V12 = V13 = V20 = V21 = V23 = V30 = V31 = V32 = R(0);
    }
    const R lxx0 = R(2) * th.Q0, lxx1 = R(2) * th.Q1, lxx2 = R(2) * th.Q2, lxx3 = R(2) * th.qb;
    const R luu0 = R(2) * th.R0, luu1 = R(2) * th.R1;
    const R gam = th.gamma, dt = E.dt;
    struct RicIn {
        R px, py, pt, pb, u0, u1;
        RefPt<R> r;
        R gx, gy;
    };
    auto load_step = [&](int k, RicIn &q) {
        const R *xk = X + EL(k, TR, 0);
        q.px = ld_stream(xk + 0 * 32); q.py = ld_stream(xk + 1 * 32); q.pt = ld_stream(xk + 2 * 32);
        q.pb = ld_stream(xk + 3 * 32); q.u0 = ld_stream(xk + 4 * 32); q.u1 = ld_stream(xk + 5 * 32);
        if (P.use_ref) {
            const R *rk = ref_i + EL(k, RR, 0);
            q.r.x0 = ld_stream(rk + 0 * 32); q.r.x1 = ld_stream(rk + 1 * 32); q.r.x2 = ld_stream(rk + 2 * 32);
            q.r.u0 = ld_stream(rk + 3 * 32); q.r.u1 = ld_stream(rk + 4 * 32);
        } else {
            q.r.x0 = t0; q.r.x1 = t1; q.r.x2 = t2; q.r.u0 = R(0); q.r.u1 = R(0);
        }
        q.gx = ld_stream(xk + 6 * 32); q.gy = ld_stream(xk + 7 * 32);
    };
    R gnx = X[EL(T, TR, 6)], gny = X[EL(T, TR, 7)];  // g at x_T, then carried backwards
    RicIn in;
    int bad_k = -1;   // horizon index at which the backward sweep first produced a non-finite feed-forward gain
    for (int k = T - 1; k >= 0; --k) {
        RicIn nx;
        load_step(k, in);
        nx = in;
        const R px = in.px, py = in.py, pt = in.pt, pb = in.pb, u0 = in.u0, u1 = in.u1;
        const RefPt<R> r = in.r;
        Lin<R> L;
        linearize(E, th, pt, u0, gnx, gny, in.gx, in.gy, L);
        gnx = in.gx;
        gny = in.gy;
        in = nx;
        // stage derivatives, core/cost_derivs.py:58-76, 110-129
        const R lx0 = lxx0 * (px - r.x0), lx1 = lxx1 * (py - r.x1);
        const R lx2 = lxx2 * heading_err_for_derivs(th, pt, r.x2), lx3 = lxx3 * pb;
        const R lu0 = luu0 * (u0 - r.u0), lu1 = luu1 * (u1 - r.u1);

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
        {
            R *o = GK + EL(k, GR, 0);
            st_stream(o + 0 * 32, K00); st_stream(o + 1 * 32, K01); st_stream(o + 2 * 32, K02); st_stream(o + 3 * 32, K03);
            st_stream(o + 4 * 32, K10); st_stream(o + 5 * 32, K11); st_stream(o + 6 * 32, K12); st_stream(o + 7 * 32, K13);
            st_stream(o + 8 * 32, k0);
            st_stream(o + 9 * 32, k1);
        }
        if (bad_k < 0 && !finite_(k0 + k1)) bad_k = k;   // the sweep runs backwards: the first hit is where it broke
        // V_x = Q_x + K^T Q_uu k + K^T Q_u + Q_xu k   (un-regularised Q_uu, core/ddp.py:251)
        const R s0 = m::fma_(Quu01, k1, Quu00 * k0) + Qu0;
        const R s1 = m::fma_(Quu11, k1, Quu10 * k0) + Qu1;
        Vx0 = Qx0 + m::fma_(K10, s1, K00 * s0) + m::fma_(Qux10, k1, Qux00 * k0);
        Vx1 = Qx1 + m::fma_(K11, s1, K01 * s0) + m::fma_(Qux11, k1, Qux01 * k0);
        Vx2 = Qx2 + m::fma_(K12, s1, K02 * s0) + m::fma_(Qux12, k1, Qux02 * k0);
        Vx3 = Qx3 + m::fma_(K13, s1, K03 * s0) + m::fma_(Qux13, k1, Qux03 * k0);
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
    }
    // non-finite guard (the reference raises: core/ddp.py:225-254).  One test after the sweep sees everything: every
    // entry of V_xx / V_x feeds the same entry one step earlier (A = I + ...), and every gain feeds V_xx or V_x of its
    // own step, so a NaN or Inf produced anywhere is still there at k = 0.
    const R chk = (Vx0 + Vx1) + (Vx2 + Vx3) + (V00 + V11) + (V22 + V33) + (V01 + V02) + (V03 + V12) + (V13 + V23) +
                  (V10 + V20) + (V30 + V21) + (V31 + V32);
    if (!finite_(chk)) {
        P.ws.flags[i] |= 1;
        if (P.io.fail_at && P.io.fail_at[2 * i] < 0) {
            P.io.fail_at[2 * i] = P.ws.it[i];
            P.io.fail_at[2 * i + 1] = bad_k;
        }
    }
}

// Register cap of the line search: the kernel needs ~90 registers unconstrained; 72 costs ~60 B of spills per thread
// but fits 7 CTAs of 4 warps per SM instead of 5.  Same-box sweep (tube step, B = 65,536, full / paper mode):
// 96: 20.0 / 15.1 ms, 80: 19.7 / 14.4, 76: 19.8 / 14.5, 72: 19.4 / 14.6, 64: 19.5 / 14.5.
#ifndef DTMPC_FWD_MAXNREG
#define DTMPC_FWD_MAXNREG 72
#endif
#ifndef DTMPC_FWD_MAXNREG_F32
#define DTMPC_FWD_MAXNREG_F32 48   // the float instantiation takes 63 uncapped; 48 -> 10 CTAs per SM: step 8.07 -> 7.53 ms (44: 8.2, 40: 8.7)
#endif

// The per-instance parameters a rollout step needs (12 values) are the same for every candidate warp of the CTA and
// constant over the sweep: parked in shared memory and re-read each step they cost 12 LDS instead of ~24 registers.

constexpr int TH_ROWS = 16;
template <typename R> __device__ __forceinline__ void theta_park(R *s, int lane, const Theta<R> &t) {
    R *v = s + lane;
    v[0 * 32] = t.Q0; v[1 * 32] = t.Q1; v[2 * 32] = t.Q2; v[3 * 32] = t.R0; v[4 * 32] = t.R1; v[5 * 32] = t.qb;
    v[6 * 32] = t.gamma; v[7 * 32] = t.tight; v[8 * 32] = t.a_dyn; v[9 * 32] = t.c1_dyn; v[10 * 32] = t.c2_dyn;
    v[11 * 32] = t.c3_dyn; v[12 * 32] = t.a_jac; v[13 * 32] = t.jac_off; v[14 * 32] = t.c2_jac; v[15 * 32] = t.c3_jac;
}
template <typename R> __device__ __forceinline__ void theta_fetch(const R *s, int lane, Theta<R> &t) {
    const volatile R *v = s + lane;  // volatile: re-read every step, never hoisted back into registers
    t.Q0 = v[0 * 32]; t.Q1 = v[1 * 32]; t.Q2 = v[2 * 32]; t.R0 = v[3 * 32]; t.R1 = v[4 * 32]; t.qb = v[5 * 32];
    t.gamma = v[6 * 32]; t.tight = v[7 * 32]; t.a_dyn = v[8 * 32]; t.c1_dyn = v[9 * 32]; t.c2_dyn = v[10 * 32];
    t.c3_dyn = v[11 * 32];
}
template <typename R> __device__ __forceinline__ void theta_fetch_jac(const R *s, int lane, Theta<R> &t) {
    const volatile R *v = s + lane;
    t.a_jac = v[12 * 32]; t.jac_off = v[13 * 32]; t.c2_jac = v[14 * 32]; t.c3_jac = v[15 * 32];
}
constexpr int DEC_KEEP = 1 << 8, DEC_BAD = 1 << 9;

// ---------------------------------------------------------------------------------------------------------------
// Line search + selection + commit in ONE kernel (round 2).  blockDim = (32, n_eval): a CTA owns 32 slots.
//   phase 1  every warp rolls out one candidate and leaves its state at the segment boundaries in the CTA's
//            checkpoint tile (n_eval x (n_seg - 1) x 5 rows x 32 lanes; in the workspace, see DTMPC_CKPT_GLOBAL);
//   phase 2  warp 0 takes the decision for its 32 slots and reserves the survivors' slots of the next iteration with
//            ONE atomicAdd per CTA on the slot counter (the survivors of a CTA stay contiguous and in order; CTAs land
//            in the order they finish.  Results do not depend on which slot an instance occupies: every instance is
//            computed by the same instruction sequence, tests/test_gpu_graph.py);
//   phase 3  the warps re-roll the winner, one horizon segment each, from the checkpoints (bit-identical to the
//            candidate, dtmpc_math.cuh) and write the next trajectory -- or, for an instance that stops, the caller's
//            result tensors.  The gains and the old trajectory were read by this CTA a moment ago (L1 / L2).
// Against the separate fwd / commit kernels this drops one launch per iteration, the global checkpoint, decision and
// ballot arrays with their traffic, the O(#CTAs) rank scan in every commit CTA, and the latency-bound commit launch
// (a third of the issue efficiency of the line search, profiles/r02a): the commit work now runs at the occupancy of
// the line search.  The number of segments is the number of candidate warps.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FWDC_MAX_SEG = DTMPC_MAX_ALPHAS;
// 1: the segment checkpoints live in the workspace instead of shared memory (15 KB per CTA in full mode, 38 KB with six
// candidates): more of the SM's 256 KB stays L1 for the rows the candidate warps share.  Same-call A/B (B = 65,536,
// full mode): fwdc 413.8 -> 406.5 us per launch, step 14.06 -> 13.81 ms, +6 % DRAM traffic (the checkpoints of a tile
// are written and read back by the same CTA, mostly out of L2).
#ifndef DTMPC_CKPT_GLOBAL
#define DTMPC_CKPT_GLOBAL 1
#endif

template <typename R, int NOBS = -1, int UR = -1>
__global__ void __maxnreg__(sizeof(R) == 8 ? DTMPC_FWD_MAXNREG : DTMPC_FWD_MAXNREG_F32) fwdc_kernel(const __grid_constant__ SolveParams<R> P) {
    using m = M<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R *sJ = reinterpret_cast<R *>(smem_raw);                 // [MAX_ALPHAS][32] candidate costs
    R *sTh = sJ + DTMPC_MAX_ALPHAS * 32;                     // [TH_ROWS][32] parked parameters
#if DTMPC_CKPT_GLOBAL
    // checkpoints of this CTA's tile in the workspace (written and read back by the same CTA: L2-resident); the lane
    // offset of the shared-memory form is kept, the tile base differs
    R *sCk = P.ws.ckpt + (long long)blockIdx.x * ((long long)CR * (FWDC_MAX_SEG - 1) * P.n_eval) * 32;
#else
    R *sCk = sTh + TH_ROWS * 32;                             // [n_eval][n_seg - 1][CR][32] checkpoints
#endif
    __shared__ int sDec[32], sJn[32];
    const int n_act = P.ws.counts[P.parity];
    if ((int)blockIdx.x * 32 >= n_act) return;  // whole CTA
    const int lane = threadIdx.x, wy = threadIdx.y;
    const bool live = (int)blockIdx.x * 32 + lane < n_act;
    const int j = live ? blockIdx.x * 32 + lane : n_act - 1;  // idle lanes shadow the last slot (they reach the barriers)
    const long long i = P.ws.list[P.parity][j];
    const EnvT<R> &E = P.env;
    const int T = P.ocp.T;
    const int n_seg = P.n_seg, seg_len = P.seg_len;
    const bool valid = wy < P.n_eval;
    const int a_idx = P.eval_idx[valid ? wy : (P.n_eval > 0 ? P.n_eval - 1 : 0)];
    const R al = P.ocp.alphas[a_idx];
    R t0 = R(0), t1 = R(0), t2 = R(0);
    if (UR < 0 ? !P.use_ref : UR == 0) {
        const R *tp = P.io.target + i * P.io.target_stride;
        t0 = tp[0]; t1 = tp[1]; t2 = tp[2];
    }
    const R *X = P.ws.traj[P.parity] + tile_off(j, traj_rows(T));
    const R *GK = P.ws.gains + tile_off(j, gain_rows(T));
    const R *ref_i = P.ws.ref + tile_off(i, ref_rows(T));
    {
        Theta<R> th;
        load_theta(th, P.io.theta + i * P.io.theta_stride, P.ocp);
        if (wy == 0) theta_park(sTh, lane, th);
    }
    __syncthreads();

    // ---- phase 1: candidate rollout, cost only -------------------------------------------------------------------
    State<R> x{X[EL(0, TR, 0)], X[EL(0, TR, 1)], X[EL(0, TR, 2)], X[EL(0, TR, 3)], P.ws.B0[i]};
    R J = R(0);
    // Inputs of a rollout step: x_k, u_k (6 rows) and K_k, k_k (10 rows) of the slot.  The wait for them at the top of a
    // step is the largest single stall of the kernel (ncu source view, profiles/r02_fwdc: 21 % of all stall samples sit
    // on the first instructions that consume them).  Issuing the loads of step k+1 as soon as u_k is known (software
    // pipelining through registers) was measured and rejected: the 16 values in flight push the kernel far over its
    // 72-register budget (stack frame 104 -> 256 B) and the step went from 14.2 to 19.5 ms (18.3 ms at 80 registers).
    struct StepIn {
        R X0, X1, X2, X3, v0, v1, K00, K01, K02, K03, K10, K11, K12, K13, k0, k1;
    };
    auto load_in = [](const R *xp, const R *gp, StepIn &q) {
        q.X0 = xp[0 * 32]; q.X1 = xp[1 * 32]; q.X2 = xp[2 * 32]; q.X3 = xp[3 * 32];
        q.v0 = xp[4 * 32]; q.v1 = xp[5 * 32];
        q.K00 = gp[0 * 32]; q.K01 = gp[1 * 32]; q.K02 = gp[2 * 32]; q.K03 = gp[3 * 32];
        q.K10 = gp[4 * 32]; q.K11 = gp[5 * 32]; q.K12 = gp[6 * 32]; q.K13 = gp[7 * 32];
        q.k0 = gp[8 * 32]; q.k1 = gp[9 * 32];
    };
    {
        int k_ck = seg_len, seg = 0;
        const R *xp = X, *gp = GK, *rp = ref_i;
        StepIn in;
        for (int k = 0; k < T; ++k) {
            if (k == k_ck) {  // phase 3 restarts the winner from here (uniform branch)
                R *c = sCk + ((wy * (n_seg - 1) + seg) * CR) * 32 + lane;
                c[0 * 32] = x.px; c[1 * 32] = x.py; c[2 * 32] = x.th; c[3 * 32] = x.b; c[4 * 32] = x.Bc;
                k_ck += seg_len;
                ++seg;
            }
            load_in(xp, gp, in);
            RefPt<R> r;
            ref_at<R, UR>(P, rp, true, t0, t1, t2, r);
            rp += RR * 32;
            Theta<R> tk;
            tk.eps_dyn = P.ocp.eps_dyn; tk.angle_wrap = P.ocp.angle_wrap;
            theta_fetch(sTh, lane, tk);
            // du = k + K (x - X_k);  u = clamp(V_k + alpha du)   (core/ddp.py:267-271)
            const R d0 = x.px - in.X0, d1 = x.py - in.X1, d2 = x.th - in.X2, d3 = x.b - in.X3;
            const R du0 = in.k0 + m::fma_(in.K03, d3, m::fma_(in.K02, d2, m::fma_(in.K01, d1, in.K00 * d0)));
            const R du1 = in.k1 + m::fma_(in.K13, d3, m::fma_(in.K12, d2, m::fma_(in.K11, d1, in.K10 * d0)));
            R u0 = m::fma_(al, du0, in.v0), u1 = m::fma_(al, du1, in.v1);
            clamp_u(E, u0, u1);
            xp += TR * 32;
            gp += GR * 32;
            J += stage_cost(tk, x.px, x.py, x.th, x.b, u0, u1, r.x0, r.x1, r.x2, r.u0, r.u1);
            State<R> xs[1] = {x};
            const R a0[1] = {u0}, a1[1] = {u1};
            step_multi<R, 1, NOBS>(E, tk, xs, a0, a1);
            x = xs[0];
        }
        RefPt<R> r;
        ref_at<R, UR>(P, rp, false, t0, t1, t2, r);
        Theta<R> tk;   // the terminal weights are used once: reloaded here instead of held across the sweep
        tk.angle_wrap = P.ocp.angle_wrap;
        const R *p = P.io.theta + i * P.io.theta_stride;
        tk.Qf0 = p[5]; tk.Qf1 = p[6]; tk.Qf2 = p[7]; tk.qb = p[8];
        J += terminal_cost(tk, x.px, x.py, x.th, x.b, r.x0, r.x1, r.x2);
        if (valid) sJ[a_idx * 32 + lane] = J;
    }
    __syncthreads();

    // ---- phase 2: selection (strict '<' in list order, tolerance break on the accepted costs; core/ddp.py:293-305)
    if (wy == 0) {
        bool keep = false;
        int dec = 0;
        if (live) {
            const OcpT<R> &o = P.ocp;
            const R Jc = P.ws.Jcur[i];
            bool bad = (P.ws.flags[i] & 1) != 0;
            int best_c = -1;
            R best = R(0), second = m::inf_();
            for (int c = 0; c < o.n_alpha; ++c) {
                const R v = (o.alphas[c] == R(0)) ? Jc : sJ[c * 32 + lane];
                bad |= !finite_(v);
                if (best_c < 0 || v < best) {
                    if (best_c >= 0) second = m::min_(second, best);
                    best = v;
                    best_c = c;
                } else {
                    second = m::min_(second, v);
                }
            }
            int it = P.ws.it[i];
            bool done = true;
            if (!bad) {
                if (o.alphas[best_c] != R(0)) P.ws.Jcur[i] = best;
                if (P.io.alpha_idx) P.io.alpha_idx[i * o.max_iter + it] = best_c;
                if (P.io.costs) P.io.costs[i * o.max_iter + it] = best;
                if (P.io.margins) P.io.margins[i * o.max_iter + it] = second - best;
                const bool conv = it > 0 && (m::abs_(P.ws.prev[i] - best) < o.tol);
                P.ws.prev[i] = best;
                ++it;
                P.ws.it[i] = it;
                done = conv || it >= o.max_iter;
            } else {
                P.io.status[i] |= DTMPC_ST_NONFINITE << P.status_shift;
                if (P.io.fail_at && P.io.fail_at[2 * i] < 0) {   // a cost was non-finite (the backward pass reports itself)
                    P.io.fail_at[2 * i] = it;
                    P.io.fail_at[2 * i + 1] = -1;
                }
            }
            keep = !done;
            dec = best_c | (keep ? DEC_KEEP : 0) | (bad ? DEC_BAD : 0);
            if (!keep && P.io.iters) P.io.iters[i] = it;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        int base = 0;
        if (lane == 0 && mask) base = atomicAdd(&P.ws.counts[P.parity ^ 1], __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int jn = base + __popc(mask & ((1u << lane) - 1u));
        if (keep) P.ws.list[P.parity ^ 1][jn] = (int)i;
        sDec[lane] = dec;
        sJn[lane] = jn;
    }
    __syncthreads();
    if (!live || wy >= n_seg) return;

    // ---- phase 3: commit the winner, one horizon segment per warp ------------------------------------------------
    const int seg = wy;
    const int d = sDec[lane];
    const bool keep = (d & DEC_KEEP) != 0;
    const long long jn = sJn[lane];
    const R alw = (d & DEC_BAD) ? R(0) : P.ocp.alphas[d & 0xff];
    const bool copy = alw == R(0);  // the current trajectory stays
    const int k_lo = seg * seg_len, k_hi = min(T, k_lo + seg_len);
    R *Xn = P.ws.traj[P.parity ^ 1] + tile_off(jn, traj_rows(T));
    R X0 = X[EL(k_lo, TR, 0)], X1 = X[EL(k_lo, TR, 1)], X2 = X[EL(k_lo, TR, 2)], X3 = X[EL(k_lo, TR, 3)];
    State<R> xw{X0, X1, X2, X3, R(0)};
    if (seg == 0) {
        xw.Bc = P.ws.B0[i];
        if (keep) {
            Xn[EL(0, TR, 0)] = X0; Xn[EL(0, TR, 1)] = X1; Xn[EL(0, TR, 2)] = X2; Xn[EL(0, TR, 3)] = X3;
            Xn[EL(0, TR, 6)] = X[EL(0, TR, 6)]; Xn[EL(0, TR, 7)] = X[EL(0, TR, 7)];
        } else {
            st4(P.io.X + i * (4LL * (T + 1)), X0, X1, X2, X3);
            if (P.write_ref) {
                R *Ro = P.ws.ref + tile_off(i, ref_rows(T));
                Ro[EL(0, RR, 0)] = X0; Ro[EL(0, RR, 1)] = X1; Ro[EL(0, RR, 2)] = X2;
            }
        }
    } else if (!copy) {
        const R *c = sCk + ((P.eval_pos[d & 0xff] * (n_seg - 1) + (seg - 1)) * CR) * 32 + lane;
        xw.px = c[0 * 32]; xw.py = c[1 * 32]; xw.th = c[2 * 32]; xw.b = c[3 * 32]; xw.Bc = c[4 * 32];
    }
    const R *xp = X + EL(k_lo, TR, 0), *gp = GK + EL(k_lo, GR, 0);
    R *xn = Xn + EL(k_lo, TR, 0);
    // inputs of a commit step: u_k, K_k, k_k and the OLD x_{k+1} (the next step's reference state, or, with g_{k+1},
    // the copied successor)
    struct CommitIn {
        R v0, v1, K00, K01, K02, K03, K10, K11, K12, K13, k0, k1, N0, N1, N2, N3;
    };
    auto load_commit = [](const R *xp, const R *gp, CommitIn &q) {
        q.v0 = xp[4 * 32]; q.v1 = xp[5 * 32];
        q.K00 = gp[0 * 32]; q.K01 = gp[1 * 32]; q.K02 = gp[2 * 32]; q.K03 = gp[3 * 32];
        q.K10 = gp[4 * 32]; q.K11 = gp[5 * 32]; q.K12 = gp[6 * 32]; q.K13 = gp[7 * 32];
        q.k0 = gp[8 * 32]; q.k1 = gp[9 * 32];
        q.N0 = xp[(TR + 0) * 32]; q.N1 = xp[(TR + 1) * 32]; q.N2 = xp[(TR + 2) * 32]; q.N3 = xp[(TR + 3) * 32];
    };
    CommitIn ci;
    for (int k = k_lo; k < k_hi; ++k) {
        load_commit(xp, gp, ci);
        const R K00 = ci.K00, K01 = ci.K01, K02 = ci.K02, K03 = ci.K03, K10 = ci.K10, K11 = ci.K11, K12 = ci.K12, K13 = ci.K13;
        const R k0 = ci.k0, k1 = ci.k1, N0 = ci.N0, N1 = ci.N1, N2 = ci.N2, N3 = ci.N3;
        R u0 = ci.v0, u1 = ci.v1, gx, gy;
        if (!copy) {
            const R d0 = xw.px - X0, d1 = xw.py - X1, d2 = xw.th - X2, d3 = xw.b - X3;
            const R du0 = k0 + m::fma_(K03, d3, m::fma_(K02, d2, m::fma_(K01, d1, K00 * d0)));
            const R du1 = k1 + m::fma_(K13, d3, m::fma_(K12, d2, m::fma_(K11, d1, K10 * d0)));
            u0 = m::fma_(alw, du0, ci.v0);
            u1 = m::fma_(alw, du1, ci.v1);
            clamp_u(E, u0, u1);
        } else {
            gx = xp[(TR + 6) * 32];
            gy = xp[(TR + 7) * 32];
        }
        xp += TR * 32;
        gp += GR * 32;
        if (!keep) {   // a stopping instance returns the gains of its last backward pass (core/ddp.py:307)
            if (P.io.K) {
                R *Ko = P.io.K + i * (8LL * T) + 8 * k;
                st4(Ko, K00, K01, K02, K03);
                st4(Ko + 4, K10, K11, K12, K13);
            }
            if (P.io.kff) st2(P.io.kff + i * (2LL * T) + 2 * k, k0, k1);
        }
        if (!copy) {
            Theta<R> tk;
            tk.eps_dyn = P.ocp.eps_dyn; tk.eps_jac = P.ocp.eps_jac; tk.angle_wrap = P.ocp.angle_wrap;
            theta_fetch(sTh, lane, tk);
            theta_fetch_jac(sTh, lane, tk);
            step_g<R, NOBS>(E, tk, xw, u0, u1, gx, gy);
        } else {
            xw.px = N0; xw.py = N1; xw.th = N2; xw.b = N3;
        }
        if (keep) {
            st_stream(xn + 4 * 32, u0);
            st_stream(xn + 5 * 32, u1);
            st_stream(xn + (TR + 0) * 32, xw.px);
            st_stream(xn + (TR + 1) * 32, xw.py);
            st_stream(xn + (TR + 2) * 32, xw.th);
            st_stream(xn + (TR + 3) * 32, xw.b);
            st_stream(xn + (TR + 6) * 32, gx);
            st_stream(xn + (TR + 7) * 32, gy);
            xn += TR * 32;
        } else {
            st2(P.io.V + i * (2LL * T) + 2 * k, u0, u1);
            st4(P.io.X + i * (4LL * (T + 1)) + 4 * (k + 1), xw.px, xw.py, xw.th, xw.b);
            if (P.write_ref) {
                R *Ro = P.ws.ref + tile_off(i, ref_rows(T));
                Ro[EL(k + 1, RR, 0)] = xw.px;
                Ro[EL(k + 1, RR, 1)] = xw.py;
                Ro[EL(k + 1, RR, 2)] = xw.th;
                Ro[EL(k, RR, 3)] = u0;
                Ro[EL(k, RR, 4)] = u1;
            }
        }
        X0 = N0; X1 = N1; X2 = N2; X3 = N3;
    }
}
__host__ inline size_t fwdc_smem_bytes(int n_eval, int n_seg, size_t elem) {
    const int ne = n_eval > 0 ? n_eval : 1;
#if DTMPC_CKPT_GLOBAL
    (void)ne; (void)n_seg;
    return elem * 32 * (size_t)(DTMPC_MAX_ALPHAS + TH_ROWS);
#else
    return elem * 32 * (size_t)(DTMPC_MAX_ALPHAS + TH_ROWS + ne * (n_seg > 1 ? n_seg - 1 : 0) * CR);
#endif
}


}  // namespace dtmpc
