/*
 * dtmpc_oracle.h -- CPU restatement of the reference's tube-MPC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA product
 * in differentiable_tube_mpc_b200/.  Only tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py may load it.  The
 * product never links, imports or calls anything in oracle/.
 *
 * Every function follows the reference's dense, unbatched arithmetic in the
 * reference's own evaluation order (sequential dot products, partial-pivot LU
 * for the 2x2 solves) and cites the file:line under /root/reference it
 * restates.  Parity pin: the reference ships no golden vectors, so
 * the oracle is pinned against outputs of the reference itself, executed in
 * the build container by oracle/make_golden.py and committed under
 * tests/golden/ (checked by tests/test_oracle_golden.py).
 */
#ifndef DTMPC_ORACLE_H
#define DTMPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_OBS 16
#define ORC_MAX_ALPHAS 8

/* obstacle aggregation: configs/dubins.yaml:60, core/tube_mpc.py:690-704 */
enum { ORC_AGG_SMOOTHMIN = 0, ORC_AGG_MIN = 1, ORC_AGG_SINGLE = 2 };

/* status codes (reference raises instead: core/ddp.py:142,159,298) */
enum { ORC_OK = 0, ORC_NONFINITE = 1, ORC_NO_CANDIDATE = 2 };

typedef struct {
    double dt;                 /* core/systems/dubins.py:12 */
    double u_min[2], u_max[2]; /* core/control.py:56-57 */
    double active_tol;         /* core/control.py:58 (1e-8) */
    int n_obs;
    int obs_agg;
    double obs_cx[ORC_MAX_OBS], obs_cy[ORC_MAX_OBS], obs_r[ORC_MAX_OBS];
    double obs_beta;           /* core/systems/dubins_obstacles.py:41 */
} orc_env;

typedef struct {
    int T;            /* ILQRConfig.horizon, core/ddp.py:14 */
    int max_iter;     /* core/ddp.py:17 */
    int n_alpha;
    int angle_wrap;   /* run_nominal.py:297-324 cost variant */
    double tol, reg;  /* core/ddp.py:18-19 */
    double alphas[ORC_MAX_ALPHAS]; /* core/ddp.py:20 */
    double eps_dyn;   /* DBaSConfig.eps seen by f_hat (Q5: 1e-6 in full mode) */
    double eps_jac;   /* DBaSConfig.eps seen by f_jac / gradient closures     */
    int tight_in_jac; /* reference: 0 (dubins_aug_jac.py:103-106 ignores tightening) */
    int sens_reg_zero;/* unused, keeps layout explicit */
    /* constrained parameter values for this instance */
    double Q[3], R[2], Qf[3], qb, alpha, gamma, tight;
    double target[3];
} orc_ocp;

/* primitives ------------------------------------------------------------ */
double orc_h(const orc_env *env, const double x[3]);
void orc_grad_h(const orc_env *env, const double x[3], double g[3]);
double orc_B(double z, double alpha, double eps);
double orc_dB(double z, double alpha, double eps);
void orc_dubins_step(const orc_env *env, const double x[3], const double u[2], double xn[3]);
void orc_fhat(const orc_env *env, const orc_ocp *p, const double xh[4], const double u[2], double out[4]);
double orc_b0(const orc_env *env, const orc_ocp *p, const double x[3]);
void orc_aug_jac(const orc_env *env, const orc_ocp *p, const double xh[4], const double u[2],
                 double A[16], double Bm[8]);
void orc_clamp(const orc_env *env, const double u[2], double out[2]);
void orc_active(const orc_env *env, const double u[2], int act[2]);
double orc_stage_cost(const orc_ocp *p, const double xh[4], const double u[2],
                      const double xref[3], const double uref[2]);
double orc_terminal_cost(const orc_ocp *p, const double xh[4], const double xref[3]);

/* solver ---------------------------------------------------------------- */
/* Xref [T+1,3] / Uref [T,2] may be NULL: then the constant p->target and zero
 * controls are tracked (nominal OCP).  K [T,2,4], kff [T,2] receive the gains
 * of the LAST backward pass (core/ddp.py:210-211).  alpha_idx/costs/margins
 * have max_iter entries; margins[it] = gap between best and second-best
 * candidate cost (decision margin). */
int orc_ilqr_solve(const orc_env *env, const orc_ocp *p, const double x0[4], const double *V_init,
                   const double *Xref, const double *Uref, double *X, double *V, double *K, double *kff,
                   int *iters, int *alpha_idx, double *costs, double *margins);

/* gx [T+1,4], gu [T,2] upper-level gradients (core/ddp.py:326-328).
 * outputs dX [T+1,4], dV [T,2], dL [T+1,4], active [T,2]. */
int orc_sensitivity(const orc_env *env, const orc_ocp *p, const double *X, const double *V, const double *gx,
                    const double *gu, double *dX, double *dV, double *dL, int *active);

/* closed-form IFT accumulation (core/ift.py:35-92 differentiated by hand,
 * SURVEY App. A.10).  g12 = d total / d (Q[3],R[2],Qf[3],qb,alpha,gamma,tight)
 * w.r.t. the CONSTRAINED values; gXref [T+1,3], gUref [T,2] may be NULL. */
void orc_ift_accumulate(const orc_env *env, const orc_ocp *p, const double *X, const double *V, const double *Xref,
                        const double *Uref, const double *dX, const double *dV, const double *dL, double g12[12],
                        double *gXref, double *gUref);

/* One full two-layer solve for a batch (serial; callers thread over slices).
 * mode 0 = paper (aux sensitivity + gradient), 1 = full (+ nominal
 * sensitivity driven by gXref/gUref and theta_bar gradient).
 * All arrays are instance-major (AoS): x0n/x0a [B,4], Vn0/Va0 [B,T,2],
 * theta_n/theta_a [B,12] (stride 0 -> shared), target [3].            */
typedef struct {
    double *Xn, *Un, *Xa, *Ua;         /* [B,T+1,4] [B,T,2] ... */
    double *Kn, *kn, *Ka, *ka;         /* [B,T,2,4] [B,T,2]  (may be NULL) */
    double *dXa, *dVa, *dLa;           /* aux sensitivity */
    double *dXn, *dVn, *dLn;           /* nominal sensitivity (mode 1) */
    double *g_aux, *g_nom;             /* [B,12] */
    double *gXref, *gUref;             /* [B,T+1,3] [B,T,2] (mode 1) */
    double *loss;                      /* [B] */
    int *iters_n, *iters_a, *status;   /* [B] */
    int *aidx_n, *aidx_a;              /* [B,max_iter] */
    double *costs_n, *costs_a, *marg_n, *marg_a; /* [B,max_iter] */
} orc_tube_out;

int orc_tube_solve_batch(const orc_env *env, const orc_ocp *nom_cfg, const orc_ocp *aux_cfg, long B,
                         const double *x0n, const double *x0a, const double *Vn0, const double *Va0,
                         const double *theta_n, long theta_n_stride, const double *theta_a, long theta_a_stride,
                         int mode, int n_threads, orc_tube_out *out);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
