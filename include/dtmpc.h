/*
 * dtmpc.h -- C ABI of the B200-native batched tube-MPC hot path.
 *
 * Drop-in boundary for the solver/IFT layer of lmcggg/differentiable-tube-mpc
 * (reference files cited per entry point, relative to the reference root).
 * The reference is pure Python/PyTorch and exposes this path as functions that
 * take Python closures; a CUDA kernel cannot call closures, so the boundary is
 * declarative: a POD description of the Dubins + DBaS problem family (what
 * core/tube_mpc.py:680-768 extracts from the YAML) plus raw device pointers.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - arrays are instance-major, C-contiguous: X [B,T+1,4], V [B,T,2],
 *     K [B,T,2,4], k [B,T,2], theta [B,12] (or one shared row, stride 0);
 *   - the caller owns all memory; the library keeps no state between calls,
 *     allocates nothing and frees nothing; scratch comes from the caller
 *     (size query functions below);
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return 0 = launched, non-zero = DTMPC_E_* (nothing launched);
 *   - per-instance numerical failures are reported in status[] (the reference
 *     raises FloatingPointError instead, core/ddp.py:138-159);
 *   - _f64 / _f32 select the arithmetic type of every array marked `real`.
 *
 * theta layout (CONSTRAINED values, reference core/params.py:29-35):
 *   [0:3] Q   [3:5] R   [5:8] Qf   [8] q_b   [9] alpha   [10] gamma   [11] tight
 */
#ifndef DTMPC_H
#define DTMPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTMPC_ABI_VERSION 3
#define DTMPC_MAX_OBS 16
#define DTMPC_MAX_ALPHAS 8
#define DTMPC_NTHETA 12

enum { DTMPC_AGG_SMOOTHMIN = 0, DTMPC_AGG_MIN = 1, DTMPC_AGG_SINGLE = 2 };

/* return codes */
enum {
    DTMPC_OK = 0,
    DTMPC_E_BADARG = 1,     /* null pointer / bad size / unsupported option */
    DTMPC_E_WORKSPACE = 2,  /* scratch buffer too small */
    DTMPC_E_CUDA = 3        /* launch or runtime failure: see dtmpc_last_error() */
};

/* per-instance status bits */
enum {
    DTMPC_ST_OK = 0,
    DTMPC_ST_NONFINITE = 1,      /* reference: FloatingPointError (core/ddp.py:142,159) */
    DTMPC_ST_NO_CANDIDATE = 2,   /* reference: RuntimeError (core/ddp.py:298-299); reserved: only reachable with a
                                    feasible_fn, which no reference caller passes and this ABI does not take */
    DTMPC_ST_AUX_SHIFT = 4       /* tube solve: aux status = (status >> 4) & 15        */
};

/* Environment shared by all instances: Dubins step, control box, obstacles.
 * reference: DubinsConfig core/systems/dubins.py:10-21, BoxClampControl core/control.py:38-70,
 * CircleObstacle + aggregation core/systems/dubins_obstacles.py:10-117, tube_mpc.py:689-721. */
typedef struct dtmpc_env {
    double dt;
    double u_min[2], u_max[2];
    double active_tol;              /* 1e-8, core/control.py:58 */
    int32_t n_obs;                  /* 0 -> h = 1 everywhere (dubins_obstacles.py:57-61) */
    int32_t obs_agg;                /* DTMPC_AGG_* */
    double obs_cx[DTMPC_MAX_OBS], obs_cy[DTMPC_MAX_OBS], obs_r[DTMPC_MAX_OBS];
    double obs_beta;                /* smooth-min temperature */
} dtmpc_env;

/* One optimal-control problem family: ILQRConfig (core/ddp.py:12-20) + the DBaS
 * numerical settings the closures capture (core/barrier.py:16-33). */
typedef struct dtmpc_ocp {
    int32_t horizon;                /* T */
    int32_t max_iter;
    int32_t n_alpha;
    int32_t angle_wrap;             /* 1: wrapped heading error, run_nominal.py:297-324 */
    double tol, reg;
    double alphas[DTMPC_MAX_ALPHAS];
    double eps_dyn;                 /* DBaSConfig.eps inside f_hat                      */
    double eps_jac;                 /* DBaSConfig.eps inside f_jac and gradient closures */
    int32_t tight_in_jac;           /* 0 = reference behaviour (dubins_aug_jac.py ignores tightening) */
    int32_t reserved;
} dtmpc_ocp;

/* Buffers of one batched iLQR solve.  Optional pointers may be NULL.  X, V, K, kff must be 16-byte aligned (they are
 * written with vector stores; cudaMalloc and torch allocations are). */
typedef struct dtmpc_solve_io {
    const void *x0;                 /* real [B,4]   augmented initial state (x, y, heading, b)       */
    const void *V_init;             /* real [B,T,2] warm start                                       */
    const void *theta;              /* real [B,12] or [12]                                           */
    int64_t theta_stride;           /* 12 or 0                                                       */
    const void *target;             /* real [B,3] or [3]: constant reference when Xref == NULL       */
    int64_t target_stride;          /* 3 or 0                                                        */
    const void *Xref;               /* real [B,T+1,3] optional tracking reference                    */
    const void *Uref;               /* real [B,T,2]   optional                                       */
    void *X;                        /* real [B,T+1,4] out                                            */
    void *V;                        /* real [B,T,2]   out                                            */
    void *K;                        /* real [B,T,2,4] out, optional: gains of the last backward pass */
    void *kff;                      /* real [B,T,2]   out, optional                                  */
    int32_t *iters;                 /* [B] out, optional                                             */
    int32_t *status;                /* [B] out (required)                                            */
    int32_t *alpha_idx;             /* [B,max_iter] out, optional: accepted step index per iteration */
    void *costs;                    /* real [B,max_iter] out, optional: accepted cost per iteration  */
    void *margins;                  /* real [B,max_iter] out, optional: second-best minus best cost  */
    int32_t *fail_at;               /* [B,2] out, optional: where the first non-finite value appeared, */
                                    /* as the reference prints it (core/ddp.py:138-159, 230-235):      */
                                    /* (iteration, step).  step = horizon index of the first non-finite */
                                    /* feed-forward gain of that iteration's backward pass, or -1 when  */
                                    /* a line-search / rollout cost was non-finite; (-1,-1) = clean     */
} dtmpc_solve_io;

/* Buffers of one batched sensitivity sweep + IFT accumulation.  X, V, Uref, gu, Xbar, dX, dV, dL, gUref must be
 * 16-byte aligned (state / control rows move as vectors; cudaMalloc and torch allocations are). */
typedef struct dtmpc_sens_io {
    const void *X, *V;              /* real [B,T+1,4], [B,T,2]: the lower-level optimum               */
    const void *theta;              /* real [B,12] or [12]                                           */
    int64_t theta_stride;
    const void *target;             /* constant cost reference when Xref == NULL                     */
    int64_t target_stride;
    const void *Xref, *Uref;        /* real [B,T+1,S], [B,T,2] cost reference, optional              */
    int64_t xref_row;               /* S: row length of Xref (3, or 4 to pass a nominal X directly)  */
    const void *gx, *gu;            /* real [B,T+1,4], [B,T,2] upper-level gradients; NULL gx selects */
                                    /* the built-in loss  sum|x - xbar|^2 + sum b^2 (tube_mpc.py:916) */
    const void *Xbar;               /* real [B,T+1,4] required when gx == NULL                       */
    void *dX, *dV, *dL;             /* real out (sensitivity) / in (accumulate-only call)            */
    void *g_theta;                  /* real [B,12] out, optional: d/d(constrained theta)             */
    void *gXref, *gUref;            /* real [B,T+1,3], [B,T,2] out, optional                          */
    void *loss;                     /* real [B] out, optional (built-in loss only)                   */
    int32_t *status;                /* [B] in/out, optional: DTMPC_ST_NONFINITE is OR-ed in          */
    int64_t gx_row;                 /* row length of gx: 4 (0 means 4) or 3 (no barrier-state        */
                                    /* component: lets gXref of one sweep drive the next one)        */
} dtmpc_sens_io;

/* Thread safety and internal state.  Entry points may be called from several host threads (errors are thread-local).
 * The library owns no device memory, but it does keep, per device, a small pool of non-blocking CUDA streams and
 * events (created on first use, never destroyed): batches of >= 16,384 instances are cut into up to 4 contiguous
 * chunks that run on those streams, forked from and joined to the caller's stream by events, so the call is still
 * ordered on `stream` like any other and may be captured into a CUDA graph.  The pool's events are shared by the
 * calls on a device, so the ENQUEUE of a chunked call holds a per-device lock (the kernels of different calls still
 * overlap on the GPU).  DTMPC_CHUNKS=1 in the environment disables chunking (and with it every internal stream). */
int dtmpc_abi_version(void);
const char *dtmpc_last_error(void);
/* Number of CUDA kernels this library has enqueued in this process so far (all entry points, all streams). */
uint64_t dtmpc_kernel_launches(void);

/* Scratch sizes in bytes (elem_size 8 or 4); pure functions of their arguments (no device needed).
 * n_alpha = number of line-search step sizes (the larger of the two OCPs for a tube solve). */
size_t dtmpc_solve_workspace_bytes(int32_t T, int64_t B, int32_t n_alpha, int32_t elem_size);
size_t dtmpc_sens_workspace_bytes(int32_t T, int64_t B, int32_t elem_size);
size_t dtmpc_step_workspace_bytes(int32_t T, int64_t B, int32_t n_alpha, int32_t elem_size);
size_t dtmpc_reduce_workspace_bytes(int64_t B, int32_t ncols, int32_t elem_size);

/* Evaluates the library's fp64 elementary functions on n arguments (test hook for the <= 2 ulp claim of the
 * hand-written exp / log / sincos / reciprocal that replace libdevice's on the rollout path):
 * out [n,6] = exp_nonpos(-|x|), log_ge1(1 + |x|), log_1to2(1 + frac(|x|)), sin(x), cos(x), rcp(x). */
int dtmpc_math_selftest_f64(const void *x, int64_t n, void *out, void *stream);

/* ilqr_solve for B instances -- reference core/ddp.py:102-307 (clamp, rollout, <= max_iter x
 * [linearise :172-205, backward Riccati :208-254, all-alpha forward line search :256-301, tol break :303]). */
int dtmpc_ilqr_solve_f64(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_solve_io *io,
                         void *workspace, size_t workspace_bytes, void *stream);
int dtmpc_ilqr_solve_f32(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_solve_io *io,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Two-layer tube MPC in one call: nominal solve, then the auxiliary solve tracking the nominal
 * solution of the same instance -- reference core/tube_mpc.py:844-909 (paper) / :305-399 (full).
 * aux->Xref/Uref are ignored. status = nominal bits | auxiliary bits << 4. */
int dtmpc_tube_solve_f64(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,
                         const dtmpc_solve_io *nom_io, const dtmpc_solve_io *aux_io, void *workspace,
                         size_t workspace_bytes, void *stream);
int dtmpc_tube_solve_f32(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,
                         const dtmpc_solve_io *nom_io, const dtmpc_solve_io *aux_io, void *workspace,
                         size_t workspace_bytes, void *stream);

/* One closed-loop step of the reference in one call (core/tube_mpc.py:844-976 paper mode, :305-602 full mode):
 * nominal solve, auxiliary solve, auxiliary sensitivity + gradients (sens_aux), and -- when sens_nom is not NULL --
 * the nominal sensitivity + gradients.  The sens_io buffers are the caller's as for dtmpc_sensitivity_*: sens_aux->X/V
 * = aux_io->X/V, Xref/Uref/Xbar = nom_io->X/V (xref_row 4); sens_nom->X/V = nom_io->X/V, gx = sens_aux->gXref with
 * gx_row 3, gu = sens_aux->gUref.  Every chunk of the batch runs its whole chain on its own stream, so the
 * sensitivity sweeps of one chunk overlap the solves of the others.  workspace: dtmpc_step_workspace_bytes. */
int dtmpc_tube_step_f64(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,
                        const dtmpc_solve_io *nom_io, const dtmpc_solve_io *aux_io, const dtmpc_sens_io *sens_aux,
                        const dtmpc_sens_io *sens_nom, void *workspace, size_t workspace_bytes, void *stream);
int dtmpc_tube_step_f32(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,
                        const dtmpc_solve_io *nom_io, const dtmpc_solve_io *aux_io, const dtmpc_sens_io *sens_aux,
                        const dtmpc_sens_io *sens_nom, void *workspace, size_t workspace_bytes, void *stream);

/* ddp_sensitivity (reference core/ddp.py:317-427) fused with the IFT accumulation
 * (reference core/ift.py:35-92, closed forms; paper mode core/tube_mpc.py:965-976).
 * Precondition: X is the rollout of V through f_hat (X[k+1] = f_hat(X[k], V[k]), as every solve of this library
 * returns it and as the reference's own call sites pass it).  The kernels take h and grad h of the successor state
 * from the stored X[k+1]; the reference recomputes f(X[k], V[k]), which differs only for an inconsistent (X, V).
 * One of Xref / target is required (the cost reference of the sweep). */
int dtmpc_sensitivity_f64(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_sens_io *io,
                          void *workspace, size_t workspace_bytes, void *stream);
int dtmpc_sensitivity_f32(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_sens_io *io,
                          void *workspace, size_t workspace_bytes, void *stream);

/* ift_gradient alone, for caller-supplied dX, dV, dL (reference core/ift.py:35-92). */
int dtmpc_ift_accumulate_f64(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_sens_io *io,
                             void *stream);
int dtmpc_ift_accumulate_f32(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_sens_io *io,
                             void *stream);

/* Deterministic column sums out[c] = sum_i g[i,c] (batch-summed adaptation gradient / loss, Algorithm 2:
 * reference core/tube_mpc.py:232-249, 979-984 consume the sum).  Fixed reduction tree: bitwise reproducible. */
int dtmpc_reduce_columns_f64(const void *g, int64_t B, int32_t ncols, void *out, void *workspace,
                             size_t workspace_bytes, void *stream);
int dtmpc_reduce_columns_f32(const void *g, int64_t B, int32_t ncols, void *out, void *workspace,
                             size_t workspace_bytes, void *stream);

/* rollout (reference core/ddp.py:89-99) through f_hat = dbas_step(dubins_step) (core/barrier.py:75-108). */
int dtmpc_rollout_f64(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const void *x0, const void *V,
                      const void *theta, int64_t theta_stride, void *X, void *stream);
int dtmpc_rollout_f32(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const void *x0, const void *V,
                      const void *theta, int64_t theta_stride, void *X, void *stream);

/* Point evaluations of the L1 primitives for B (x_hat, u) pairs; any output may be NULL.
 * x_next [B,4]: f_hat (core/barrier.py:93-108); A [B,4,4], Bm [B,4,2]: dubins_augmented_jacobian
 * (core/systems/dubins_aug_jac.py:61-139); h [B], grad_h [B,3]: dubins_obstacles.py:41-117;
 * cost [B]: stage cost w.r.t. target (core/tube_mpc.py:823-827); u_clamped [B,2], active [B,2] (int32):
 * core/control.py:61-70; b0 [B]: dbas_init_b0 = B(h(x) - tight) with the jacobian-time eps (core/barrier.py:111-120). */
int dtmpc_primitives_f64(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const void *xh, const void *u,
                         const void *theta, int64_t theta_stride, const void *target, void *x_next, void *A,
                         void *Bm, void *h, void *grad_h, void *cost, void *u_clamped, int32_t *active,
                         void *b0, void *stream);
int dtmpc_primitives_f32(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const void *xh, const void *u,
                         const void *theta, int64_t theta_stride, const void *target, void *x_next, void *A,
                         void *Bm, void *h, void *grad_h, void *cost, void *u_clamped, int32_t *active,
                         void *b0, void *stream);

/* Disturbance drawn inside the plant step when no w tensor is passed (reference core/systems/dubins.py:57-65:
 * w = low + (high - low) * U[0,1)^3).  Philox4x32-10, key = seed, counter = (first_instance + i, *step, draw): a pure
 * function of (seed, global instance, step), hence independent of batch size, sharding and CUDA-graph replay. */
typedef struct dtmpc_disturbance {
    double low[3], high[3];
    uint64_t seed;
    int64_t first_instance;         /* global index of instance 0 of this call (offset of the rank's shard)     */
    const int64_t *step;            /* DEVICE counter t of the closed-loop step, read at kernel run time        */
} dtmpc_disturbance;

/* Closed-loop plant step for B tubes (reference core/tube_mpc.py:990-1023): x+ = f(x,u0)+w, b+ from the
 * undisturbed successor, nominal propagation, warm-start shift.  In-place on xh/xh_bar/V*_ws.
 * w [B,3] (device) or NULL; with w == NULL and rng != NULL the disturbance is generated in the kernel. */
int dtmpc_plant_step_f64(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,
                         const void *Un, const void *Ua, const void *theta_n, int64_t theta_n_stride,
                         const void *theta_a, int64_t theta_a_stride, const void *w, const dtmpc_disturbance *rng,
                         void *xh, void *xh_bar, void *Vn_ws, void *Va_ws, void *stream);
int dtmpc_plant_step_f32(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,
                         const void *Un, const void *Ua, const void *theta_n, int64_t theta_n_stride,
                         const void *theta_a, int64_t theta_a_stride, const void *w, const dtmpc_disturbance *rng,
                         void *xh, void *xh_bar, void *Vn_ws, void *Va_ws, void *stream);

/* FP64 FMA peak of the current device (dependent-chain-free DFMA loop), TFLOP/s; used by bench.py as the
 * roofline denominator because MEASURED_PEAKS.json has no FP64 figure.  Synchronous. */
double dtmpc_measure_fp64_peak_tflops(int32_t iters, void *stream);
double dtmpc_measure_fp32_peak_tflops(int32_t iters, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DTMPC_H */
