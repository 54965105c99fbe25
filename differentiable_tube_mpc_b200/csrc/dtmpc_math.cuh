// dtmpc_math.cuh -- device primitives of the Dubins + DBaS problem family.
//
// Compiled with -fmad=false: every fused multiply-add below is written
// explicitly (fma_), so a formula produces the same bits wherever it is
// inlined (rollout, line-search candidates, winner recomputation, jacobian).
// The line search relies on that: the alpha = 0 candidate is the current
// trajectory bit for bit and is never re-evaluated.
//
// Reference formulas (file:line relative to the reference root) are cited per
// function; evaluation order follows them except where a dense product with
// structural zeros/ones is written out sparsely (adding exact zeros).
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/dtmpc.h"

#include "dtmpc_fastmath_coeffs.inc"

#ifndef DTMPC_SELECT_MINMAX
#define DTMPC_SELECT_MINMAX 1
#endif
#ifndef DTMPC_FAST_RCP
#define DTMPC_FAST_RCP 1     // reciprocals by MUFU seed + two Newton steps instead of the IEEE division (see fm::rcp_d)
#endif
#ifndef DTMPC_FAST_GUARDS
#define DTMPC_FAST_GUARDS 1  // exp of a non-positive difference / log of a sum >= 1 without the general-range guards:
                             // -3 % on the final kernels (16.1 -> 15.6 ms per step); it cost +9 % on the first design
#endif

namespace dtmpc {

// ---------------------------------------------------------------------------
// fp64 exp / log / sincos with their polynomial coefficients in __constant__ memory.
// libdevice carries fp64 coefficients as 64-bit immediates; sm_100a has no 64-bit immediate operand, so every use
// costs two UMOV + two MOV on top of the DFMA (measured: 236 UMOV + 160 MOV of 1750 issued instructions per
// line-search step, profiles/r01b).  A constant-bank operand is free.  Accuracy: <= 2 ulp on the argument ranges
// the path uses (tools/gen_fastmath.py prints the fit errors); outside them the functions defer to libdevice.
// ---------------------------------------------------------------------------
namespace fm {
__constant__ double EXP_C[DTMPC_EXP_N] = {DTMPC_EXP_COEFFS};
__constant__ double LOG_C[DTMPC_LOG_N] = {DTMPC_LOG_COEFFS};
__constant__ double SIN_C[DTMPC_SIN_N] = {DTMPC_SIN_COEFFS};
__constant__ double COS_C[DTMPC_COS_N] = {DTMPC_COS_COEFFS};
__constant__ double K[12] = {DTMPC_LOG2E, -DTMPC_LN2_HI, -DTMPC_LN2_LO, DTMPC_LN2_HI, DTMPC_LN2_LO, DTMPC_2OPI,
                             -DTMPC_PIO2_1, -DTMPC_PIO2_2, -DTMPC_PIO2_3, 0x1.8p52, 0x1.6a09e667f3bcdp+0, 0.0};
enum { LOG2E = 0, NLN2_HI, NLN2_LO, LN2_HI, LN2_LO, TWO_OPI, NPIO2_1, NPIO2_2, NPIO2_3, MAGIC, SQRT2 };

// 1/x from the hardware seed (MUFU.RCP64H: the reciprocal of the upper 32 bits of x) and two Newton steps: relative
// error 2^-20 -> 2^-40 -> 2^-80, i.e. the correctly rounded reciprocal or its neighbour (<= 1 ulp), in 1 + 4
// instructions and without the slow-path call of the IEEE division (~14 instructions, a branch and a subroutine per use;
// profiles/r01g: 3 divisions per line-search step).  Operands on this path are normal numbers (sums >= 1, barrier
// arguments >= eps, pivots of Q_uu + reg I); 0, subnormals and NaN give inf / NaN, which the non-finite guards report
// exactly as they report the IEEE results inf / NaN.
__device__ __forceinline__ double rcp_d(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = ::fma(-x, r, 1.0);
    r = ::fma(r, e, r);
    e = ::fma(-x, r, 1.0);
    return ::fma(r, e, r);
}

// exp(x) for x <= ~709; x < -700 flushes to 0 (the path only exponentiates z_i - z_max <= 0)
__device__ __forceinline__ double exp_d(double x) {
    const double t = ::fma(x, K[LOG2E], K[MAGIC]);
    const int k = __double2loint(t);
    const double kf = t - K[MAGIC];
    double r = ::fma(kf, K[NLN2_HI], x);
    r = ::fma(kf, K[NLN2_LO], r);
    double p = EXP_C[0];
#pragma unroll
    for (int j = 1; j < DTMPC_EXP_N; ++j) p = ::fma(p, r, EXP_C[j]);
    double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    if (x < -700.0) res = 0.0;
    if (!(x <= 709.0)) res = ::exp(x);  // overflow range, NaN: never on the hot path
    return res;
}

// exp(x) for x <= 0 (differences z_i - z_max of the smooth-min): no overflow guard.  The argument is clamped at -700:
// e^-700 = 1e-304 cannot change a sum >= 1 (nor the gradient sums weighted by it) in fp64, which is all the path does
// with the result.  NaN propagates: arithmetic NaNs are canonical (low word 0), so k = 0 and p = NaN survives.
__device__ __forceinline__ double exp_nonpos_d(double x) {
    x = x < -700.0 ? -700.0 : x;
    const double t = ::fma(x, K[LOG2E], K[MAGIC]);
    const int k = __double2loint(t);
    const double kf = t - K[MAGIC];
    double r = ::fma(kf, K[NLN2_HI], x);
    r = ::fma(kf, K[NLN2_LO], r);
    double p = EXP_C[0];
#pragma unroll
    for (int j = 1; j < DTMPC_EXP_N; ++j) p = ::fma(p, r, EXP_C[j]);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// log(s) for s >= 1 finite (s = sum of exp(<= 0) with one term equal to 1); NaN propagates
__device__ __forceinline__ double log_ge1_d(double s) {
    const int hi = __double2hiint(s);
    int e = (hi >> 20) - 1023;
    double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(s));
    const bool big = m > K[SQRT2];
    m = big ? m * 0.5 : m;
    e = big ? e + 1 : e;
#if DTMPC_FAST_RCP
    const double u = (m - 1.0) * rcp_d(m + 1.0);
#else
    const double u = (m - 1.0) / (m + 1.0);
#endif
    const double w = u * u;
    double p = LOG_C[0];
#pragma unroll
    for (int j = 1; j < DTMPC_LOG_N; ++j) p = ::fma(p, w, LOG_C[j]);
    const double u2 = u + u;
    const double lm = ::fma(u2 * w, p, u2);
    const double ef = (double)e;
    const double res = ::fma(ef, K[LN2_HI], lm) + ef * K[LN2_LO];
    return s != s ? s : res;
}

// log(s) for 1 <= s <= 2 (the two-term smooth-min: s = 1 + exp(z2 - z1)).  Same operations, hence the same bits, as
// log_ge1_d on that range, but the exponent is known (0, or 1 after the sqrt(2) fold), so there is no bit surgery and
// a NaN propagates through the arithmetic by itself.
__device__ __forceinline__ double log_1to2_d(double s) {
    const bool big = s > K[SQRT2];
    const double m = big ? s * 0.5 : s;
#if DTMPC_FAST_RCP
    const double u = (m - 1.0) * rcp_d(m + 1.0);
#else
    const double u = (m - 1.0) / (m + 1.0);
#endif
    const double w = u * u;
    double p = LOG_C[0];
#pragma unroll
    for (int j = 1; j < DTMPC_LOG_N; ++j) p = ::fma(p, w, LOG_C[j]);
    const double u2 = u + u;
    const double lm = ::fma(u2 * w, p, u2);
    return big ? (K[LN2_HI] + lm) + K[LN2_LO] : lm;
}

// log(s) for positive normal s (the path calls it with s = sum of exp(<= 0) in [1, n_obs])
__device__ __forceinline__ double log_d(double s) {
    if (!(s >= 2.2250738585072014e-308 && s < __longlong_as_double(0x7ff0000000000000LL))) return ::log(s);
    const int hi = __double2hiint(s);
    int e = (hi >> 20) - 1023;
    double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(s));
    if (m > K[SQRT2]) {
        m *= 0.5;
        e += 1;
    }
    const double u = (m - 1.0) / (m + 1.0);
    const double w = u * u;
    double p = LOG_C[0];
#pragma unroll
    for (int j = 1; j < DTMPC_LOG_N; ++j) p = ::fma(p, w, LOG_C[j]);
    const double u2 = u + u;
    const double lm = ::fma(u2 * w, p, u2);  // 2 atanh(u)
    const double ef = (double)e;
    return ::fma(ef, K[LN2_HI], lm) + ef * K[LN2_LO];
}

__device__ __forceinline__ void sincos_d(double x, double *sp, double *cp) {
    if (!(::fabs(x) < 1.0e5)) {  // large arguments, inf, NaN: libdevice (Payne-Hanek)
        ::sincos(x, sp, cp);
        return;
    }
    const double t = ::fma(x, K[TWO_OPI], K[MAGIC]);
    const int q = __double2loint(t);
    const double qf = t - K[MAGIC];
    double r = ::fma(qf, K[NPIO2_1], x);
    r = ::fma(qf, K[NPIO2_2], r);
    r = ::fma(qf, K[NPIO2_3], r);
    const double w = r * r;
    double ps = SIN_C[0], pc = COS_C[0];
#pragma unroll
    for (int j = 1; j < DTMPC_SIN_N; ++j) ps = ::fma(ps, w, SIN_C[j]);
#pragma unroll
    for (int j = 1; j < DTMPC_COS_N; ++j) pc = ::fma(pc, w, COS_C[j]);
    const double sn = ::fma(r * w, ps, r);
    const double cs = ::fma(w * w, pc, ::fma(-0.5, w, 1.0));
    const double a = (q & 1) ? cs : sn;
    const double b = (q & 1) ? sn : cs;
    *sp = (q & 2) ? -a : a;
    *cp = ((q + 1) & 2) ? -b : b;
}
}  // namespace fm

template <typename R> struct M;  // math dispatch
template <> struct M<double> {
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return ::fma(a, b, c); }
#ifdef DTMPC_LIBDEVICE_MATH
    static __device__ __forceinline__ double exp_(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log_(double x) { return ::log(x); }
    static __device__ __forceinline__ double exp_nonpos_(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log_ge1_(double x) { return ::log(x); }
    static __device__ __forceinline__ double log_1to2_(double x) { return ::log(x); }
    static __device__ __forceinline__ void sincos_(double x, double *s, double *c) { ::sincos(x, s, c); }
#else
    static __device__ __forceinline__ double exp_(double x) { return fm::exp_d(x); }
    static __device__ __forceinline__ double log_(double x) { return fm::log_d(x); }
#if DTMPC_FAST_GUARDS
    static __device__ __forceinline__ double exp_nonpos_(double x) { return fm::exp_nonpos_d(x); }
    static __device__ __forceinline__ double log_ge1_(double x) { return fm::log_ge1_d(x); }
    static __device__ __forceinline__ double log_1to2_(double x) { return fm::log_1to2_d(x); }
#else
    static __device__ __forceinline__ double exp_nonpos_(double x) { return fm::exp_d(x); }
    static __device__ __forceinline__ double log_ge1_(double x) { return fm::log_d(x); }
    static __device__ __forceinline__ double log_1to2_(double x) { return fm::log_d(x); }
#endif
    static __device__ __forceinline__ void sincos_(double x, double *s, double *c) { fm::sincos_d(x, s, c); }
#endif
    static __device__ __forceinline__ double atan2_(double y, double x) { return ::atan2(y, x); }
#if DTMPC_FAST_RCP
    static __device__ __forceinline__ double rcp_(double x) { return fm::rcp_d(x); }
#else
    static __device__ __forceinline__ double rcp_(double x) { return 1.0 / x; }
#endif
    // compare-select (DSETP + 2 FSEL) instead of fmax/fmin (~9 instructions with their NaN rules); a NaN in `a`
    // propagates, which is what torch.clamp / torch.maximum do
#if DTMPC_SELECT_MINMAX
    static __device__ __forceinline__ double max_(double a, double b) { return a < b ? b : a; }
    static __device__ __forceinline__ double min_(double a, double b) { return a > b ? b : a; }
#else
    static __device__ __forceinline__ double max_(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ double min_(double a, double b) { return ::fmin(a, b); }
#endif
    static __device__ __forceinline__ double abs_(double a) { return ::fabs(a); }
    static __device__ __forceinline__ double inf_() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double nan_() { return __longlong_as_double(0x7ff8000000000000LL); }
};
template <> struct M<float> {
    static __device__ __forceinline__ float fma_(float a, float b, float c) { return ::fmaf(a, b, c); }
    static __device__ __forceinline__ float exp_(float x) { return ::expf(x); }
    static __device__ __forceinline__ float log_(float x) { return ::logf(x); }
    static __device__ __forceinline__ float exp_nonpos_(float x) { return ::expf(x); }
    static __device__ __forceinline__ float log_ge1_(float x) { return ::logf(x); }
    static __device__ __forceinline__ float log_1to2_(float x) { return ::logf(x); }
    static __device__ __forceinline__ float rcp_(float x) { return 1.0f / x; }
    static __device__ __forceinline__ void sincos_(float x, float *s, float *c) { ::sincosf(x, s, c); }
    static __device__ __forceinline__ float atan2_(float y, float x) { return ::atan2f(y, x); }
    static __device__ __forceinline__ float max_(float a, float b) { return a < b ? b : a; }
    static __device__ __forceinline__ float min_(float a, float b) { return a > b ? b : a; }
    static __device__ __forceinline__ float abs_(float a) { return ::fabsf(a); }
    static __device__ __forceinline__ float inf_() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float nan_() { return __int_as_float(0x7fc00000); }
};

template <typename R> __device__ __forceinline__ bool finite_(R v) { return M<R>::abs_(v) < M<R>::inf_(); }

// ---------------------------------------------------------------------------
// Kernel-parameter copies of the POD specs, converted to the arithmetic type
// on the host exactly as the reference converts Python floats to tensors.
// ---------------------------------------------------------------------------
template <typename R> struct EnvT {
    R dt, ndt;                  // dt and -dt
    R umin0, umin1, umax0, umax1;
    R act_lo0, act_lo1, act_hi0, act_hi1;  // u_min + tol, u_max - tol (core/control.py:70)
    R nbeta, ninv_beta;         // -beta, -(1/beta)
    int n_obs, agg;
    R cx[DTMPC_MAX_OBS], cy[DTMPC_MAX_OBS], r2[DTMPC_MAX_OBS];
};

template <typename R> struct OcpT {
    int T, max_iter, n_alpha, angle_wrap, tight_in_jac;
    R tol, reg;
    R alphas[DTMPC_MAX_ALPHAS];
    R eps_dyn, eps_jac;
};

// Per-instance constrained parameters + derived barrier constants.
template <typename R> struct Theta {
    R Q0, Q1, Q2, R0, R1, Qf0, Qf1, Qf2, qb, alpha, gamma, tight;
    // a = max(alpha, eps) for the dynamics (eps_dyn) and for the jacobian/gradients (eps_jac)
    R a_dyn, a_jac, jac_off;
    R c1_dyn, c2_dyn, c3_dyn;  // 1/a, 1/a^2, 1/a^3 of the relaxed branch (a = a_dyn)
    R c2_jac, c3_jac;          // 1/a^2, 1/a^3 with a = a_jac: relaxed branch of B' (barrier_dB_jac)
    // per-OCP numerical settings, kept with the lane so that lanes of one warp may work on
    // different OCPs (nominal / auxiliary) without divergent constant-bank reads
    R eps_dyn, eps_jac;
    int angle_wrap;
};

template <typename R>
__device__ __forceinline__ void load_theta(Theta<R> &t, const R *__restrict__ p, const OcpT<R> &o) {
    t.Q0 = p[0]; t.Q1 = p[1]; t.Q2 = p[2];
    t.R0 = p[3]; t.R1 = p[4];
    t.Qf0 = p[5]; t.Qf1 = p[6]; t.Qf2 = p[7];
    t.qb = p[8]; t.alpha = p[9]; t.gamma = p[10]; t.tight = p[11];
    t.eps_dyn = o.eps_dyn;
    t.eps_jac = o.eps_jac;
    t.angle_wrap = o.angle_wrap;
    t.a_dyn = M<R>::max_(t.alpha, o.eps_dyn);  // core/barrier.py:49
    t.a_jac = M<R>::max_(t.alpha, o.eps_jac);
    t.jac_off = o.tight_in_jac ? t.tight : R(0);
    t.c1_dyn = R(1) / t.a_dyn;
    t.c2_dyn = R(1) / (t.a_dyn * t.a_dyn);
    t.c3_dyn = R(1) / (t.a_dyn * t.a_dyn * t.a_dyn);
    t.c2_jac = R(1) / (t.a_jac * t.a_jac);
    t.c3_jac = R(1) / (t.a_jac * t.a_jac * t.a_jac);
}

// ---------------------------------------------------------------------------
// Safety function h and its gradient.
//   h_i = dx*dx + dy*dy - r^2                      core/systems/dubins_obstacles.py:27-30
//   smooth-min: z = -beta*h_i, lse = zmax + log(sum exp(z - zmax)), h = -(1/beta)*lse   :63-69
//   grad: softmax(z) weighted sum of 2(p - c_i)      :84-92   (normalised once by the sum)
//   min: min_i h_i, gradient of the first argmin     :104-106, 115-117
//   single circle: h_0, 2(p - c_0)                   :16-38
//   empty list: h = 1, grad = 0                      :57-61, 81-82
// ---------------------------------------------------------------------------
// NOBS > 0 fixes the obstacle count at compile time for a smooth-min environment (the hot kernels are instantiated
// for the counts in DTMPC_FIXED_NOBS_LIST): the obstacle loop unrolls, its first three insertions lose the compares
// against -inf, and the aggregation-mode branches disappear.  The arithmetic per point is unchanged.
template <typename R, int I> __device__ __forceinline__ void top3_insert(R z, R &z1, R &z2, R &z3, bool &p1, bool &p2) {
    if (I == 0) {  // z1 = z2 = z3 = -inf before the first obstacle
        p1 = true; p2 = true;
        z1 = z;
    } else if (I == 1) {
        p1 = z > z1; p2 = true;
        z2 = p1 ? z1 : z;
        z1 = p1 ? z : z1;
    } else if (I == 2) {
        p1 = z > z1; p2 = z > z2;
        z3 = p2 ? z2 : z;
        z2 = p1 ? z1 : (p2 ? z : z2);
        z1 = p1 ? z : z1;
    } else {
        p1 = z > z1; p2 = z > z2;
        const bool p3 = z > z3;
        z3 = p2 ? z2 : (p3 ? z : z3);
        z2 = p1 ? z1 : (p2 ? z : z2);
        z1 = p1 ? z : z1;
    }
}
template <int I, int N> struct ObsLoop {  // compile-time obstacle index
    template <typename F> __device__ __forceinline__ static void run(F &&f) {
        f(std::integral_constant<int, I>{});
        ObsLoop<I + 1, N>::run(f);
    }
};
template <int N> struct ObsLoop<N, N> {
    template <typename F> __device__ __forceinline__ static void run(F &&) {}
};

template <typename R, bool GRAD, int NOBS = -1>
__device__ __forceinline__ void eval_h(const EnvT<R> &e, R px, R py, R &h, R &gx, R &gy) {
    using m = M<R>;
    constexpr bool FIX = NOBS > 0;
    const int n = FIX ? NOBS : e.n_obs;
    gx = R(0);
    gy = R(0);
    if (!FIX && n == 0) {
        h = R(1);
        return;
    }
    if (!FIX && e.agg == DTMPC_AGG_SINGLE) {
        R dx = px - e.cx[0], dy = py - e.cy[0];
        h = m::fma_(dy, dy, dx * dx) - e.r2[0];
        if (GRAD) {
            gx = R(2) * dx;
            gy = R(2) * dy;
        }
        return;
    }
    if (!FIX && e.agg == DTMPC_AGG_MIN) {
        R best = m::inf_();
        R bx = R(0), by = R(0);
        for (int i = 0; i < n; ++i) {
            R dx = px - e.cx[i], dy = py - e.cy[i];
            R hi = m::fma_(dy, dy, dx * dx) - e.r2[i];
            if (hi < best || i == 0) {
                best = hi;
                bx = dx;
                by = dy;
            }
        }
        h = best;
        if (GRAD) {
            gx = R(2) * bx;
            gy = R(2) * by;
        }
        return;
    }
    // Smooth-min.  One pass finds the two largest z_i = -beta h_i (with their 2(p - c_i)) and the third largest
    // value.  exp(z_i - z_1) <= e^-40 = 4e-18 cannot change sum >= 1 in fp64, so when the third largest is out of
    // reach (almost always: it needs three obstacles within 2/beta of each other in h) the log-sum-exp has two terms
    // and costs ONE exp instead of n; otherwise the full ordered sum of the reference is evaluated.  The choice
    // depends only on this lane's own point, so h stays a pure function of its argument (bitwise reproducible
    // whatever the neighbouring lanes do).
    R z1 = -m::inf_(), z2 = -m::inf_(), z3 = -m::inf_();
    R ax1 = R(0), ay1 = R(0), ax2 = R(0), ay2 = R(0);
    if (FIX) {
        ObsLoop<0, FIX ? NOBS : 0>::run([&](auto I) {
            constexpr int i = decltype(I)::value;
            const R dx = px - e.cx[i], dy = py - e.cy[i];
            const R hi = m::fma_(dy, dy, dx * dx) - e.r2[i];
            const R z = e.nbeta * hi;
            bool p1, p2;
            top3_insert<R, i>(z, z1, z2, z3, p1, p2);
            if (GRAD) {
                if (i == 0) {
                    ax1 = dx; ay1 = dy;
                } else {
                    ax2 = p1 ? ax1 : (p2 ? dx : ax2);
                    ay2 = p1 ? ay1 : (p2 ? dy : ay2);
                    ax1 = p1 ? dx : ax1;
                    ay1 = p1 ? dy : ay1;
                }
            }
        });
    } else {
        for (int i = 0; i < n; ++i) {
            const R dx = px - e.cx[i], dy = py - e.cy[i];
            const R hi = m::fma_(dy, dy, dx * dx) - e.r2[i];
            const R z = e.nbeta * hi;
            const bool p1 = z > z1, p2 = z > z2, p3 = z > z3;
            z3 = p2 ? z2 : (p3 ? z : z3);
            z2 = p1 ? z1 : (p2 ? z : z2);
            z1 = p1 ? z : z1;
            if (GRAD) {
                ax2 = p1 ? ax1 : (p2 ? dx : ax2);
                ay2 = p1 ? ay1 : (p2 ? dy : ay2);
                ax1 = p1 ? dx : ax1;
                ay1 = p1 ? dy : ay1;
            }
        }
    }
    const R reach = sizeof(R) == 8 ? R(-40) : R(-18);
    R s, sx = R(0), sy = R(0);
    const bool far = z3 - z1 > reach;
    if (far) {
        s = R(0);
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const R dx = px - e.cx[i], dy = py - e.cy[i];
            const R hi = m::fma_(dy, dy, dx * dx) - e.r2[i];
            const R w = m::exp_nonpos_(e.nbeta * hi - z1);
            s += w;
            if (GRAD) {
                sx = m::fma_(w, R(2) * dx, sx);
                sy = m::fma_(w, R(2) * dy, sy);
            }
        }
    } else {
        const R e2 = m::exp_nonpos_(z2 - z1);  // 0 for a single obstacle (z2 = -inf)
        s = R(1) + e2;
        if (GRAD) {
            sx = m::fma_(e2, R(2) * ax2, R(2) * ax1);
            sy = m::fma_(e2, R(2) * ay2, R(2) * ay1);
        }
    }
    h = e.ninv_beta * (z1 + (far ? m::log_ge1_(s) : m::log_1to2_(s)));
    if (GRAD) {
        const R inv_s = m::rcp_(s);
        gx = sx * inv_s;
        gy = sy * inv_s;
    }
}

// Relaxed inverse barrier, core/barrier.py:36-59, with a = max(alpha, eps) precomputed.
template <typename R> __device__ __forceinline__ R barrier_B(R z, R a, R eps) {
    if (z >= a) return R(1) / M<R>::max_(z, eps);
    R d = z - a;
    return (R(1) / a) - d / (a * a) + (d * d) / (a * a * a);
}

// d B / d zeta, core/systems/dubins_aug_jac.py:26-40
template <typename R> __device__ __forceinline__ R barrier_dB(R z, R a, R eps) {
    if (z >= a) {
        R zc = M<R>::max_(z, eps);
        return R(-1) / (zc * zc);
    }
    R d = z - a;
    return -(R(1) / (a * a)) + (R(2) * d) / (a * a * a);
}

// d B / d a in the relaxed branch (0 in the safe branch); differentiates core/barrier.py:56-58.
template <typename R> __device__ __forceinline__ R barrier_dB_da(R z, R a) {
    if (z >= a) return R(0);
    R d = z - a;
    R a2 = a * a;
    return R(-3) * d * d / (a2 * a2);
}

// core/control.py:61-64
template <typename R> __device__ __forceinline__ void clamp_u(const EnvT<R> &e, R &u0, R &u1) {
    u0 = M<R>::min_(M<R>::max_(u0, e.umin0), e.umax0);
    u1 = M<R>::min_(M<R>::max_(u1, e.umin1), e.umax1);
}

// core/control.py:66-70
template <typename R> __device__ __forceinline__ void active_u(const EnvT<R> &e, R u0, R u1, bool &a0, bool &a1) {
    a0 = (u0 <= e.act_lo0) || (u0 >= e.act_hi0);
    a1 = (u1 <= e.act_lo1) || (u1 >= e.act_hi1);
}

// run_nominal.py:32-34
template <typename R> __device__ __forceinline__ R wrap_angle(R err) {
    R s, c;
    M<R>::sincos_(err, &s, &c);
    return M<R>::atan2_(s, c);
}

// h (value only) at NC points at once.  The obstacle loop is shared, so the NC evaluations form one instruction
// stream the scheduler can interleave (the line search rolls NC candidates per thread for instruction-level
// parallelism: a lone fp64 dependency chain issues once per ~9 cycles).  Per point the arithmetic is exactly that of
// eval_h<R, false>; every rollout in the library goes through this function, so results are bit-identical wherever a
// state is propagated (initial rollout, candidates, commit, plant step).
template <typename R, int NC, int NOBS = -1>
__device__ __forceinline__ void eval_h_multi(const EnvT<R> &e, const R (&px)[NC], const R (&py)[NC], R (&h)[NC]) {
    using m = M<R>;
    constexpr bool FIX = NOBS > 0;
    const int n = FIX ? NOBS : e.n_obs;
    if (!FIX && n == 0) {
#pragma unroll
        for (int q = 0; q < NC; ++q) h[q] = R(1);
        return;
    }
    if (!FIX && e.agg == DTMPC_AGG_SINGLE) {
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const R dx = px[q] - e.cx[0], dy = py[q] - e.cy[0];
            h[q] = m::fma_(dy, dy, dx * dx) - e.r2[0];
        }
        return;
    }
    if (!FIX && e.agg == DTMPC_AGG_MIN) {
        R best[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) best[q] = m::inf_();
        for (int i = 0; i < n; ++i) {
            const R cx = e.cx[i], cy = e.cy[i], r2 = e.r2[i];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                const R dx = px[q] - cx, dy = py[q] - cy;
                const R hi = m::fma_(dy, dy, dx * dx) - r2;
                if (hi < best[q] || i == 0) best[q] = hi;
            }
        }
#pragma unroll
        for (int q = 0; q < NC; ++q) h[q] = best[q];
        return;
    }
    R z1[NC], z2[NC], z3[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) z1[q] = z2[q] = z3[q] = -m::inf_();
    if (FIX) {
        ObsLoop<0, FIX ? NOBS : 0>::run([&](auto I) {
            constexpr int i = decltype(I)::value;
            const R cx = e.cx[i], cy = e.cy[i], r2 = e.r2[i];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                const R dx = px[q] - cx, dy = py[q] - cy;
                const R hi = m::fma_(dy, dy, dx * dx) - r2;
                bool p1, p2;
                top3_insert<R, i>(e.nbeta * hi, z1[q], z2[q], z3[q], p1, p2);
            }
        });
    } else {
        for (int i = 0; i < n; ++i) {
            const R cx = e.cx[i], cy = e.cy[i], r2 = e.r2[i];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                const R dx = px[q] - cx, dy = py[q] - cy;
                const R hi = m::fma_(dy, dy, dx * dx) - r2;
                const R z = e.nbeta * hi;
                const bool p1 = z > z1[q], p2 = z > z2[q], p3 = z > z3[q];
                z3[q] = p2 ? z2[q] : (p3 ? z : z3[q]);
                z2[q] = p1 ? z1[q] : (p2 ? z : z2[q]);
                z1[q] = p1 ? z : z1[q];
            }
        }
    }
    const R reach = sizeof(R) == 8 ? R(-40) : R(-18);
    R lg[NC];
    bool far = false;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        lg[q] = m::log_1to2_(R(1) + m::exp_nonpos_(z2[q] - z1[q]));  // exp(-inf) adds nothing for a single obstacle
        far |= (z3[q] - z1[q] > reach);
    }
    if (far) {  // three or more obstacles within reach of one of the points: the reference's full ordered sum (rare)
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            if (z3[q] - z1[q] > reach) {
                R acc = R(0);
#pragma unroll 1
                for (int i = 0; i < n; ++i) {
                    const R dx = px[q] - e.cx[i], dy = py[q] - e.cy[i];
                    const R hi = m::fma_(dy, dy, dx * dx) - e.r2[i];
                    acc += m::exp_nonpos_(e.nbeta * hi - z1[q]);
                }
                lg[q] = m::log_ge1_(acc);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NC; ++q) h[q] = e.ninv_beta * (z1[q] + lg[q]);
}

// Relaxed inverse barrier of the dynamics (core/barrier.py:36-59) without a data-dependent branch: the relaxed
// polynomial 1/a - d/a^2 + d^2/a^3 uses the per-instance reciprocals precomputed in Theta.
template <typename R> __device__ __forceinline__ R barrier_B_dyn(const Theta<R> &t, R z) {
    using m = M<R>;
    const R safe = m::rcp_(m::max_(z, t.eps_dyn));
    const R d = z - t.a_dyn;
    const R relaxed = m::fma_(d * d, t.c3_dyn, m::fma_(-d, t.c2_dyn, t.c1_dyn));
    return z >= t.a_dyn ? safe : relaxed;
}

// B' of the jacobian / gradient closures (core/systems/dubins_aug_jac.py:26-40, a = a_jac, eps = eps_jac) without a
// data-dependent branch: -1/max(z, eps)^2 in the safe set, -1/a^2 + 2 (z - a)/a^3 below it (reciprocals from Theta).
template <typename R> __device__ __forceinline__ R barrier_dB_jac(const Theta<R> &t, R z) {
    using m = M<R>;
    const R zc = m::max_(z, t.eps_jac);
    const R safe = -m::rcp_(zc * zc);
    const R relaxed = m::fma_(R(2) * (z - t.a_jac), t.c3_jac, -t.c2_jac);
    return z >= t.a_jac ? safe : relaxed;
}

// Augmented state carried along a rollout: (px, py, th, b) plus B(h_bar(x)) of the CURRENT state, which the next
// DBaS step needs as B_curr (core/barrier.py:94-98); h(x_{k+1}) of step k is h(x_k) of step k+1, so it is evaluated once.
template <typename R> struct State {
    R px, py, th, b, Bc;
};

// B_curr for a state that did not come out of step() (the initial state).
template <typename R>
__device__ __forceinline__ R barrier_of_state(const EnvT<R> &e, const Theta<R> &t, R px, R py) {
    R pxa[1] = {px}, pya[1] = {py}, h[1];
    eval_h_multi<R, 1>(e, pxa, pya, h);
    return barrier_B_dyn(t, h[0] - t.tight);
}

// One step of f_hat for NC states: dubins_step (core/systems/dubins.py:36-42) + dbas_step (core/barrier.py:93-108).
template <typename R, int NC, int NOBS = -1>
__device__ __forceinline__ void step_multi(const EnvT<R> &e, const Theta<R> &t, State<R> (&x)[NC], const R (&u0)[NC],
                                           const R (&u1)[NC]) {
    using m = M<R>;
    R pxn[NC], pyn[NC], thn[NC], h[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        R s, c;
        m::sincos_(x[q].th, &s, &c);
        const R dtv = e.dt * u0[q];
        pxn[q] = m::fma_(dtv, c, x[q].px);
        pyn[q] = m::fma_(dtv, s, x[q].py);
        thn[q] = m::fma_(e.dt, u1[q], x[q].th);
    }
    eval_h_multi<R, NC, NOBS>(e, pxn, pyn, h);
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const R Bn = barrier_B_dyn(t, h[q] - t.tight);
        // b+ = B_next - gamma * (B_curr - b)
        const R bn = m::fma_(-t.gamma, x[q].Bc - x[q].b, Bn);
        x[q].px = pxn[q];
        x[q].py = pyn[q];
        x[q].th = thn[q];
        x[q].b = bn;
        x[q].Bc = Bn;
    }
}

// step() that also returns g = B'(h_jac(x+)) grad h(x+), the quantity the barrier row of the augmented jacobian is
// built from (core/systems/dubins_aug_jac.py:95-129).  The state update is bit-identical to step(): h comes out of
// the same operation sequence in eval_h<R, true> and eval_h_multi.
template <typename R, int NOBS = -1>
__device__ __forceinline__ void step_g(const EnvT<R> &e, const Theta<R> &t, State<R> &x, R u0, R u1, R &gx, R &gy) {
    using m = M<R>;
    R s, c;
    m::sincos_(x.th, &s, &c);
    const R dtv = e.dt * u0;
    const R pxn = m::fma_(dtv, c, x.px);
    const R pyn = m::fma_(dtv, s, x.py);
    const R thn = m::fma_(e.dt, u1, x.th);
    R h, hx, hy;
    eval_h<R, true, NOBS>(e, pxn, pyn, h, hx, hy);
    const R Bn = barrier_B_dyn(t, h - t.tight);
    const R dB = barrier_dB_jac(t, h - t.jac_off);
    gx = dB * hx;
    gy = dB * hy;
    const R bn = m::fma_(-t.gamma, x.Bc - x.b, Bn);
    x.px = pxn;
    x.py = pyn;
    x.th = thn;
    x.b = bn;
    x.Bc = Bn;
}

template <typename R>
__device__ __forceinline__ void step(const EnvT<R> &e, const Theta<R> &t, State<R> &x, R u0, R u1) {
    State<R> xs[1] = {x};
    const R a0[1] = {u0}, a1[1] = {u1};
    step_multi<R, 1>(e, t, xs, a0, a1);
    x = xs[0];
}

// Stage cost  sum Q dx^2 + sum R du^2 + q_b b^2   (core/tube_mpc.py:823-827, 875-880; 281-286, 358-364)
template <typename R>
__device__ __forceinline__ R stage_cost(const Theta<R> &t, R px, R py, R th, R b, R u0, R u1, R xr0, R xr1, R xr2,
                                        R ur0, R ur1) {
    using m = M<R>;
    R d0 = px - xr0, d1 = py - xr1, d2 = th - xr2;
    if (t.angle_wrap) d2 = wrap_angle(d2);
    R e0 = u0 - ur0, e1 = u1 - ur1;
    R sx = (t.Q0 * d0) * d0;
    sx = m::fma_(t.Q1 * d1, d1, sx);
    sx = m::fma_(t.Q2 * d2, d2, sx);
    R su = (t.R0 * e0) * e0;
    su = m::fma_(t.R1 * e1, e1, su);
    return m::fma_(t.qb, b * b, sx + su);
}

// Terminal cost  sum Qf dx^2 + q_b b^2   (core/tube_mpc.py:829-832, 882-885; 288-293, 366-371)
template <typename R>
__device__ __forceinline__ R terminal_cost(const Theta<R> &t, R px, R py, R th, R b, R xr0, R xr1, R xr2) {
    using m = M<R>;
    R d0 = px - xr0, d1 = py - xr1, d2 = th - xr2;
    if (t.angle_wrap) d2 = wrap_angle(d2);
    R sx = (t.Qf0 * d0) * d0;
    sx = m::fma_(t.Qf1 * d1, d1, sx);
    sx = m::fma_(t.Qf2 * d2, d2, sx);
    return m::fma_(t.qb, b * b, sx);
}

// Heading error used by the cost DERIVATIVES: x - target with the target shifted so that the
// error is the wrapped one (run_nominal.py:311-315); plain difference otherwise.
template <typename R> __device__ __forceinline__ R heading_err_for_derivs(const Theta<R> &t, R th, R xr2) {
    if (!t.angle_wrap) return th - xr2;
    R tk = th - wrap_angle(th - xr2);
    return th - tk;
}

// g(x) = B'(h_jac(x)) * grad h(x): the barrier row of the augmented jacobian is built from g at
// x_k and x_{k+1} (core/systems/dubins_aug_jac.py:95-129).  Also returns h(x) (unshifted).
template <typename R>
__device__ __forceinline__ void barrier_grad(const EnvT<R> &e, const Theta<R> &t, R px, R py, R &gx, R &gy, R &h) {
    R hx, hy;
    eval_h<R, true>(e, px, py, h, hx, hy);
    R dB = barrier_dB_jac(t, h - t.jac_off);
    gx = dB * hx;
    gy = dB * hy;
}

// Non-trivial entries of A = d f_hat / d x_hat (4x4) and B = d f_hat / d u (4x2):
//   A = [1 0 a02 0; 0 1 a12 0; 0 0 1 0; r0 r1 r2 gamma],  B = [b00 0; b10 0; 0 dt; s0 0]
// (core/systems/dubins_aug_jac.py:43-58, 128-138; B[3,1] = g_theta * dt = 0 because h has no heading term).
template <typename R> struct Lin {
    R a02, a12, r0, r1, r2, b00, b10, s0;
};

template <typename R>
__device__ __forceinline__ void linearize(const EnvT<R> &e, const Theta<R> &t, R th, R v, R gnx, R gny, R gcx, R gcy,
                                          Lin<R> &L) {
    using m = M<R>;
    R s, c;
    m::sincos_(th, &s, &c);
    L.a02 = (e.ndt * v) * s;
    L.a12 = (e.dt * v) * c;
    L.b00 = e.dt * c;
    L.b10 = e.dt * s;
    L.r0 = m::fma_(-t.gamma, gcx, gnx);
    L.r1 = m::fma_(-t.gamma, gcy, gny);
    L.r2 = m::fma_(gny, L.a12, gnx * L.a02);
    L.s0 = m::fma_(gny, L.b10, gnx * L.b00);
}

// 2x2 solve with partial pivoting (what torch.linalg.solve / LAPACK getrf+getrs does, core/ddp.py:243-244):
// factor once, then solve_rhs per right-hand side.
// The two pivots are inverted once (two divisions) and every right-hand side is solved with multiplications:
// 1 ulp away from the divide-per-entry form of getrs, 9 fp64 divisions (~20 instructions each) cheaper per step.
template <typename R> struct LU2 {
    R a01, l, inv00, inv11;
    bool swap;
};
template <typename R> __device__ __forceinline__ void lu2_factor(LU2<R> &f, R q00, R q01, R q10, R q11) {
    f.swap = M<R>::abs_(q10) > M<R>::abs_(q00);
    R r00 = f.swap ? q10 : q00, r01 = f.swap ? q11 : q01;
    R r10 = f.swap ? q00 : q10, r11 = f.swap ? q01 : q11;
    f.a01 = r01;
    f.inv00 = M<R>::rcp_(r00);
    f.l = r10 * f.inv00;
    f.inv11 = M<R>::rcp_(M<R>::fma_(-f.l, r01, r11));
}
template <typename R> __device__ __forceinline__ void lu2_solve(const LU2<R> &f, R b0, R b1, R &x0, R &x1) {
    R c0 = f.swap ? b1 : b0, c1 = f.swap ? b0 : b1;
    R y1 = M<R>::fma_(-f.l, c0, c1);
    x1 = y1 * f.inv11;
    x0 = M<R>::fma_(-f.a01, x1, c0) * f.inv00;
}

}  // namespace dtmpc
