// dtmpc.cu -- C ABI (include/dtmpc.h) over the sm_100a kernels.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
// (see __graft_entry__.build()).  No torch types cross this boundary.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <mutex>

#include "dtmpc_math.cuh"
#include "dtmpc_sens.cuh"
#include "dtmpc_solve.cuh"

using namespace dtmpc;

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_kernel_launches{0};  // every kernel this library enqueues (dtmpc_kernel_launches)
inline void count_launches(unsigned long long n) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

int fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int cuda_fail(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return DTMPC_E_CUDA;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;  // B200
    }
    return cached;
}

// Python floats enter reference tensors by a plain cast to the tensor dtype; do the same here.
template <typename R> void convert_env(const dtmpc_env &e, EnvT<R> &o) {
    o.dt = (R)e.dt;
    o.ndt = (R)(-e.dt);
    o.umin0 = (R)e.u_min[0]; o.umin1 = (R)e.u_min[1];
    o.umax0 = (R)e.u_max[0]; o.umax1 = (R)e.u_max[1];
    o.act_lo0 = o.umin0 + (R)e.active_tol; o.act_lo1 = o.umin1 + (R)e.active_tol;
    o.act_hi0 = o.umax0 - (R)e.active_tol; o.act_hi1 = o.umax1 - (R)e.active_tol;
    o.nbeta = (R)(-e.obs_beta);
    o.ninv_beta = (R)(-(1.0 / e.obs_beta));
    o.n_obs = e.n_obs;
    o.agg = e.obs_agg;
    for (int i = 0; i < DTMPC_MAX_OBS; ++i) {
        o.cx[i] = (R)e.obs_cx[i];
        o.cy[i] = (R)e.obs_cy[i];
        o.r2[i] = (R)(e.obs_r[i] * e.obs_r[i]);
    }
}
template <typename R> void convert_ocp(const dtmpc_ocp &p, OcpT<R> &o) {
    o.T = p.horizon;
    o.max_iter = p.max_iter;
    o.n_alpha = p.n_alpha;
    o.angle_wrap = p.angle_wrap;
    o.tight_in_jac = p.tight_in_jac;
    o.tol = (R)p.tol;
    o.reg = (R)p.reg;
    for (int i = 0; i < DTMPC_MAX_ALPHAS; ++i) o.alphas[i] = (R)p.alphas[i];
    o.eps_dyn = (R)p.eps_dyn;
    o.eps_jac = (R)p.eps_jac;
}

bool env_ok(const dtmpc_env *e) {
    return e && e->n_obs >= 0 && e->n_obs <= DTMPC_MAX_OBS && e->obs_agg >= 0 && e->obs_agg <= 2 && e->dt > 0.0 &&
           (e->n_obs == 0 || e->obs_agg != DTMPC_AGG_SMOOTHMIN || e->obs_beta > 0.0);
}
bool ocp_ok(const dtmpc_ocp *p) {
    return p && p->horizon >= 1 && p->max_iter >= 1 && p->n_alpha >= 1 && p->n_alpha <= DTMPC_MAX_ALPHAS;
}

template <typename R> void convert_io(const dtmpc_solve_io &s, SolveArgs<R> &a) {
    a.x0 = (const R *)s.x0; a.V_init = (const R *)s.V_init; a.theta = (const R *)s.theta;
    a.target = (const R *)s.target; a.Xref = (const R *)s.Xref; a.Uref = (const R *)s.Uref;
    a.theta_stride = s.theta_stride; a.target_stride = s.target_stride;
    a.X = (R *)s.X; a.V = (R *)s.V; a.K = (R *)s.K; a.kff = (R *)s.kff;
    a.iters = s.iters; a.status = s.status; a.alpha_idx = s.alpha_idx;
    a.costs = (R *)s.costs; a.margins = (R *)s.margins;
    a.fail_at = s.fail_at;
}

// ---------------------------------------------------------------------------
// iLQR solver: workspace carving and the per-iteration launch sequence
// ---------------------------------------------------------------------------
size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// Carves the caller's scratch into the solver arrays; with base == nullptr only the size is computed.
template <typename R> size_t carve_solve_ws(char *base, int T, long long B, [[maybe_unused]] int n_alpha, SolveWs<R> *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off = align_up(off + bytes);
        return p;
    };
    const size_t nB = (size_t)B, nT = (size_t)((B + 31) / 32) * 32;  // tile-major arrays cover whole tiles
    char *counts = take(256);
    char *l0 = take(nB * sizeof(int)), *l1 = take(nB * sizeof(int));
    char *it = take(nB * sizeof(int)), *flags = take(nB * sizeof(int));
    char *Jcur = take(nB * sizeof(R)), *prev = take(nB * sizeof(R)), *B0 = take(nB * sizeof(R));
    char *traj0 = take(nT * sizeof(R) * (size_t)traj_rows(T)), *traj1 = take(nT * sizeof(R) * (size_t)traj_rows(T));
    char *gains = take(nT * sizeof(R) * (size_t)gain_rows(T));
    char *ref = take(nT * sizeof(R) * (size_t)ref_rows(T));
#if DTMPC_CKPT_GLOBAL
    char *ckpt = take(nT * sizeof(R) * CR * (size_t)(FWDC_MAX_SEG - 1) * (size_t)n_alpha);
#endif
    if (w) {
        w->counts = (int *)counts;
        w->list[0] = (int *)l0; w->list[1] = (int *)l1;
        w->it = (int *)it; w->flags = (int *)flags;
        w->Jcur = (R *)Jcur; w->prev = (R *)prev; w->B0 = (R *)B0;
        w->traj[0] = (R *)traj0; w->traj[1] = (R *)traj1; w->gains = (R *)gains; w->ref = (R *)ref;
#if DTMPC_CKPT_GLOBAL
        w->ckpt = (R *)ckpt;
#else
        w->ckpt = nullptr;
#endif
    }
    return off;
}

bool io_ok(const dtmpc_solve_io *s, bool need_ref_or_target) {
    if (!s || !s->x0 || !s->V_init || !s->theta || !s->X || !s->V) return false;
    if (need_ref_or_target && !s->target && !(s->Xref && s->Uref)) return false;
    if ((s->Xref == nullptr) != (s->Uref == nullptr)) return false;
    return true;
}

// One batched OCP: init, then max_iter x (backward Riccati sweep; line search + selection + commit).  Everything is
// stream-ordered; the host never waits, so all max_iter iterations are enqueued and the ones that find the slots all
// vacated return at once.
template <typename R> int launch_ocp(SolveParams<R> &P, cudaStream_t st) {
    const long long B = P.B;
    const int T = P.ocp.T;
    const unsigned gB = (unsigned)((B + LS_BLOCK - 1) / LS_BLOCK);
    P.n_eval = 0;
    for (int c = 0; c < P.ocp.n_alpha; ++c)
        if (P.ocp.alphas[c] != R(0)) P.eval_idx[P.n_eval++] = c;
    for (int c = 0, q = 0; c < P.ocp.n_alpha; ++c) P.eval_pos[c] = P.ocp.alphas[c] != R(0) ? q++ : -1;
    P.n_seg = P.n_eval > 0 ? P.n_eval : 1;   // one horizon segment per candidate warp (fwdc_kernel, phase 3)
    if (P.n_seg > T) P.n_seg = T;
    P.seg_len = (T + P.n_seg - 1) / P.n_seg;
    P.n_seg = (T + P.seg_len - 1) / P.seg_len;
    // kernels with the obstacle count fixed at compile time (1..8 circles, smooth-min; the reference scenario has 5)
    static const bool allow_fixed = !getenv("DTMPC_NO_FIXED_NOBS");
    const int fixed_nobs = (allow_fixed && P.env.agg == DTMPC_AGG_SMOOTHMIN) ? P.env.n_obs : -1;
    P.parity = 0;
#define DTMPC_INIT_CASE(N) case N: init_kernel<R, N><<<gB, LS_BLOCK, 0, st>>>(P); break;
    switch (fixed_nobs) {
        DTMPC_INIT_CASE(1) DTMPC_INIT_CASE(2) DTMPC_INIT_CASE(3) DTMPC_INIT_CASE(4)
        DTMPC_INIT_CASE(5) DTMPC_INIT_CASE(6) DTMPC_INIT_CASE(7) DTMPC_INIT_CASE(8)
        default: init_kernel<R><<<gB, LS_BLOCK, 0, st>>>(P);
    }
#undef DTMPC_INIT_CASE
    const unsigned gFwd = (unsigned)((B + 31) / 32);
    const unsigned gRic = (unsigned)((B + RIC_BLOCK - 1) / RIC_BLOCK);
    const size_t smem = fwdc_smem_bytes(P.n_eval, P.n_seg, sizeof(R));
    const dim3 blk(32, P.n_eval > 0 ? P.n_eval : 1);
    for (int it = 0; it < P.ocp.max_iter; ++it) {   // the last iteration retires every instance still running
        ric_kernel<R><<<gRic, RIC_BLOCK, 0, st>>>(P);
#define DTMPC_FWDC_LAUNCH(N, U)                                                                          \
    {                                                                                                    \
        static bool big_smem_ok = false;   /* > 48 KB of dynamic shared memory needs the opt-in, once */ \
        if (smem > 48 * 1024 && !big_smem_ok) {                                                          \
            cudaFuncSetAttribute(fwdc_kernel<R, N, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
            big_smem_ok = true;                                                                          \
        }                                                                                                \
        fwdc_kernel<R, N, U><<<gFwd, blk, smem, st>>>(P);                                                \
    }
#define DTMPC_FWDC_CASE(N)                                                \
    case N:                                                               \
        if (P.use_ref) DTMPC_FWDC_LAUNCH(N, 1) else DTMPC_FWDC_LAUNCH(N, 0) \
        break;
        switch (fixed_nobs) {
            DTMPC_FWDC_CASE(1) DTMPC_FWDC_CASE(2) DTMPC_FWDC_CASE(3) DTMPC_FWDC_CASE(4)
            DTMPC_FWDC_CASE(5) DTMPC_FWDC_CASE(6) DTMPC_FWDC_CASE(7) DTMPC_FWDC_CASE(8)
            default: DTMPC_FWDC_LAUNCH(-1, -1)
        }
#undef DTMPC_FWDC_CASE
#undef DTMPC_FWDC_LAUNCH
        P.parity ^= 1;
    }
    count_launches(1ull + 2ull * (unsigned long long)P.ocp.max_iter);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "iLQR kernel launch");
    return DTMPC_OK;
}

template <typename R> void convert_sens(const dtmpc_sens_io &s, SensArgs<R> &a) {
    a.X = (const R *)s.X; a.V = (const R *)s.V; a.theta = (const R *)s.theta; a.target = (const R *)s.target;
    a.Xref = (const R *)s.Xref; a.Uref = (const R *)s.Uref; a.gx = (const R *)s.gx; a.gu = (const R *)s.gu;
    a.Xbar = (const R *)s.Xbar;
    a.theta_stride = s.theta_stride; a.target_stride = s.target_stride; a.xref_row = s.xref_row;
    a.gx_row = s.gx_row == 0 ? 4 : s.gx_row;
    a.dX = (R *)s.dX; a.dV = (R *)s.dV; a.dL = (R *)s.dL; a.g_theta = (R *)s.g_theta;
    a.gXref = (R *)s.gXref; a.gUref = (R *)s.gUref; a.loss = (R *)s.loss; a.status = s.status;
}

bool sens_io_ok(const dtmpc_sens_io *s, bool accumulate_only) {
    if (!s || !s->X || !s->V || !s->theta || !s->dX || !s->dV || !s->dL) return false;
    if ((s->Xref == nullptr) != (s->Uref == nullptr)) return false;
    if (s->Xref && s->xref_row != 3 && s->xref_row != 4) return false;
    if (s->gx_row != 0 && s->gx_row != 3 && s->gx_row != 4) return false;
    if (!s->Xref && !s->target) return false;   // both sweeps read the cost reference (also for the deltas alone)
    for (const void *q : {s->X, s->V, s->Uref, s->gu, s->Xbar, (const void *)s->dX, (const void *)s->dV, (const void *)s->dL,
                          (const void *)s->gUref})   // state / control rows move as 16-byte vectors
        if (((uintptr_t)q & 15u) != 0) return false;
    if (!accumulate_only && !s->gx && !s->Xbar) return false;
    return true;
}

template <typename R> void offset_sens_io(SensArgs<R> &a, long long off, int T) {
    a.X += off * 4LL * (T + 1);
    a.V += off * 2LL * T;
    a.theta += off * a.theta_stride;
    if (a.target) a.target += off * a.target_stride;
    if (a.Xref) a.Xref += off * a.xref_row * (T + 1);
    if (a.Uref) a.Uref += off * 2LL * T;
    if (a.gx) a.gx += off * a.gx_row * (T + 1);
    if (a.gu) a.gu += off * 2LL * T;
    if (a.Xbar) a.Xbar += off * 4LL * (T + 1);
    a.dX += off * 4LL * (T + 1);
    a.dV += off * 2LL * T;
    a.dL += off * 4LL * (T + 1);
    if (a.g_theta) a.g_theta += off * DTMPC_NTHETA;
    if (a.gXref) a.gXref += off * 3LL * (T + 1);
    if (a.gUref) a.gUref += off * 2LL * T;
    if (a.loss) a.loss += off;
    if (a.status) a.status += off;
}

template <typename R> void launch_sens_kernel(const SensParams<R> &P, unsigned grid, cudaStream_t st) {
    static const bool allow_fixed = !getenv("DTMPC_NO_FIXED_NOBS");
#define DTMPC_SENS_CASE(N) case N: sens_kernel<R, N><<<grid, SENS_BLOCK, 0, st>>>(P); break;
    switch ((allow_fixed && P.env.agg == DTMPC_AGG_SMOOTHMIN) ? P.env.n_obs : -1) {
        DTMPC_SENS_CASE(1) DTMPC_SENS_CASE(2) DTMPC_SENS_CASE(3) DTMPC_SENS_CASE(4)
        DTMPC_SENS_CASE(5) DTMPC_SENS_CASE(6) DTMPC_SENS_CASE(7) DTMPC_SENS_CASE(8)
        default: sens_kernel<R><<<grid, SENS_BLOCK, 0, st>>>(P);
    }
#undef DTMPC_SENS_CASE
}

// sensitivity sweep of instances [off, off + Bc) on stream st, scratch at ws (sens_ws_rows(T) rows of whole tiles)
template <typename R>
int launch_sens_chunk(const dtmpc_env *env, const dtmpc_ocp *ocp, const dtmpc_sens_io *io, long long off, long long Bc,
                      void *ws, cudaStream_t st) {
    SensParams<R> P;
    memset(&P, 0, sizeof(P));
    convert_env(*env, P.env);
    convert_ocp(*ocp, P.ocp);
    convert_sens(*io, P.io);
    offset_sens_io(P.io, off, ocp->horizon);
    P.B = Bc;
    P.ws = (R *)ws;
    launch_sens_kernel<R>(P, (unsigned)((Bc + SENS_BLOCK - 1) / SENS_BLOCK), st);
    count_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "sensitivity kernel launch");
    return DTMPC_OK;
}

// ---------------------------------------------------------------------------
// Chunked execution.  The backward and commit sweeps are serial chains of T steps with one thread per instance: at
// B = 65,536 they are a single wave of 14 warps per SM whose duration is the latency of one chain however few
// instances are left, while the line search is bound by the fp64 pipe.  Contiguous chunks of the batch are
// independent problems, so they run on their own streams and the sweeps of one chunk overlap the line search of
// another.  Results do not depend on the chunking (every instance is computed by the same instruction sequence).
// ---------------------------------------------------------------------------
#ifndef DTMPC_DEFAULT_CHUNKS
#define DTMPC_DEFAULT_CHUNKS 4
#endif
constexpr int MAX_CHUNKS = 8;
constexpr int MAX_DEVICES = 64;
struct ChunkStreams {
    cudaStream_t s[MAX_CHUNKS];
    cudaEvent_t fork, done[MAX_CHUNKS];
    std::mutex enqueue;   // the events above are shared by every chunked call on this device (f32 and f64 alike)
    bool ready = false;
};
ChunkStreams *chunk_streams() {
    static ChunkStreams pool[MAX_DEVICES];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    ChunkStreams &c = pool[dev];
    if (!c.ready) {
        if (cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int i = 0; i < MAX_CHUNKS; ++i) {
            if (cudaStreamCreateWithFlags(&c.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&c.done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        c.ready = true;
    }
    return &c;
}
int solve_chunks(long long B) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("DTMPC_CHUNKS");
        forced = e ? atoi(e) : 0;
    }
    int n = forced > 0 ? forced : DTMPC_DEFAULT_CHUNKS;
    if (n > MAX_CHUNKS) n = MAX_CHUNKS;
    while (n > 1 && B / n < 8192) --n;  // a chunk should still fill the machine for the line search
    return n < 1 ? 1 : n;
}
long long chunk_size(long long B, int n) {
    const long long per = (B + n - 1) / n;
    return (per + LS_BLOCK - 1) / LS_BLOCK * LS_BLOCK;
}
size_t chunked_ws_bytes_slack() { return (size_t)MAX_CHUNKS * 24 * 256; }

template <typename R> void offset_io(SolveArgs<R> &a, long long off, int T, int max_iter) {
    a.x0 += off * 4;
    a.V_init += off * 2LL * T;
    a.theta += off * a.theta_stride;
    if (a.target) a.target += off * a.target_stride;
    if (a.Xref) a.Xref += off * 3LL * (T + 1);
    if (a.Uref) a.Uref += off * 2LL * T;
    a.X += off * 4LL * (T + 1);
    a.V += off * 2LL * T;
    if (a.K) a.K += off * 8LL * T;
    if (a.kff) a.kff += off * 2LL * T;
    if (a.iters) a.iters += off;
    if (a.status) a.status += off;
    if (a.alpha_idx) a.alpha_idx += off * max_iter;
    if (a.costs) a.costs += off * max_iter;
    if (a.margins) a.margins += off * max_iter;
    if (a.fail_at) a.fail_at += off * 2;
}

template <typename R>
int solve_impl(const dtmpc_env *env, const dtmpc_ocp *o0, const dtmpc_ocp *o1, long long B, const dtmpc_solve_io *io0,
               const dtmpc_solve_io *io1, void *ws, size_t ws_bytes, void *stream, const dtmpc_sens_io *sa = nullptr,
               const dtmpc_sens_io *sn = nullptr, bool with_sens = false) {
    if (!env_ok(env) || !ocp_ok(o0) || (o1 && !ocp_ok(o1))) return fail(DTMPC_E_BADARG, "bad env/ocp description");
    if (o1 && o1->horizon != o0->horizon) return fail(DTMPC_E_BADARG, "nominal and auxiliary horizons differ");
    if (B < 0 || B > 0x7fffffffLL) return fail(DTMPC_E_BADARG, "batch size out of range");
    if (B == 0) return DTMPC_OK;
    if (!io_ok(io0, true) || !io0->status) return fail(DTMPC_E_BADARG, "missing buffer in solve io");
    if (o1 && !io_ok(io1, false)) return fail(DTMPC_E_BADARG, "missing buffer in auxiliary solve io");
    if (!ws) return fail(DTMPC_E_BADARG, "null workspace");
    for (const dtmpc_solve_io *s : {io0, io1}) {  // results are written with 16-byte vector stores
        if (!s) continue;
        for (const void *q : {(const void *)s->X, (const void *)s->V, (const void *)s->K, (const void *)s->kff})
            if (((uintptr_t)q & 15u) != 0) return fail(DTMPC_E_BADARG, "result buffers must be 16-byte aligned");
    }
    const int n_alpha = o1 && o1->n_alpha > o0->n_alpha ? o1->n_alpha : o0->n_alpha;
    if (with_sens) {
        if (!o1 || !sens_io_ok(sa, false) || (sn && !sens_io_ok(sn, false)))
            return fail(DTMPC_E_BADARG, "missing buffer in the sensitivity io of a tube step");
        if (ws_bytes < dtmpc_step_workspace_bytes(o0->horizon, B, n_alpha, (int)sizeof(R)))
            return fail(DTMPC_E_WORKSPACE, "step workspace too small");
    } else if (ws_bytes < dtmpc_solve_workspace_bytes(o0->horizon, B, n_alpha, (int)sizeof(R))) {
        return fail(DTMPC_E_WORKSPACE, "solve workspace too small");
    }
    const cudaStream_t user = (cudaStream_t)stream;
    const int T = o0->horizon;

    int n_chunks = solve_chunks(B);
    ChunkStreams *cs = n_chunks > 1 ? chunk_streams() : nullptr;
    if (!cs) n_chunks = 1;
    const long long per = chunk_size(B, n_chunks);
    // the pool's fork/done events are shared by every call on this device: enqueue one chunked call at a time
    std::unique_lock<std::mutex> enqueue_lock;
    if (n_chunks > 1) enqueue_lock = std::unique_lock<std::mutex>(cs->enqueue);
    if (n_chunks > 1) {
        cudaError_t e = cudaEventRecord(cs->fork, user);
        if (e != cudaSuccess) return cuda_fail(e, "chunk fork");
    }
    size_t ws_off = 0;
    int rc = DTMPC_OK;
    for (int c = 0; c < n_chunks && rc == DTMPC_OK; ++c) {
        const long long off = c * per;
        if (off >= B) break;
        const long long Bc = B - off < per ? B - off : per;
        const cudaStream_t st = c == 0 ? user : cs->s[c];
        if (c > 0) {
            cudaError_t e = cudaStreamWaitEvent(st, cs->fork, 0);
            if (e != cudaSuccess) return cuda_fail(e, "chunk fork");
        }
        SolveParams<R> P;
        memset(&P, 0, sizeof(P));
        convert_env(*env, P.env);
        convert_ocp(*o0, P.ocp);
        convert_io(*io0, P.io);
        offset_io(P.io, off, T, o0->max_iter);
        ws_off += carve_solve_ws<R>((char *)ws + ws_off, T, Bc, n_alpha, &P.ws);
        P.B = Bc;
        P.use_ref = io0->Xref != nullptr;
        P.ref_from_io = P.use_ref;
        P.write_ref = o1 != nullptr;
        P.status_shift = 0;
        rc = launch_ocp(P, st);
        if (rc == DTMPC_OK && o1) {
            // auxiliary OCP tracking the nominal solution that the commit sweep left in ws.ref (core/tube_mpc.py:864-865)
            convert_ocp(*o1, P.ocp);
            convert_io(*io1, P.io);
            P.io.status = io0->status;
            offset_io(P.io, off, T, o1->max_iter);
            P.use_ref = 1;
            P.ref_from_io = 0;
            P.write_ref = 0;
            P.status_shift = 4;
            rc = launch_ocp(P, st);
        }
        if (rc == DTMPC_OK && with_sens) {
            // the chunk's sensitivity sweeps follow its solves on the same stream (scratch: after the solver arrays)
            void *sws = (char *)ws + ws_off;
            ws_off += align_up((size_t)sens_ws_rows(T) * (size_t)((Bc + 31) / 32 * 32) * sizeof(R));
            rc = launch_sens_chunk<R>(env, o1, sa, off, Bc, sws, st);
            if (rc == DTMPC_OK && sn) rc = launch_sens_chunk<R>(env, o0, sn, off, Bc, sws, st);
        }
        if (c > 0) {  // join, also after a failed launch: the caller's stream must not run ahead of the chunk
            cudaError_t e = cudaEventRecord(cs->done[c], st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(user, cs->done[c], 0);
            if (e != cudaSuccess) return cuda_fail(e, "chunk join");
        }
    }
    return rc;
}

template <typename R>
int sens_impl(const dtmpc_env *env, const dtmpc_ocp *ocp, long long B, const dtmpc_sens_io *io, void *ws, size_t ws_bytes,
              void *stream, bool accumulate_only) {
    if (!env_ok(env) || !ocp_ok(ocp)) return fail(DTMPC_E_BADARG, "bad env/ocp description");
    if (B < 0) return fail(DTMPC_E_BADARG, "negative batch");
    if (B == 0) return DTMPC_OK;
    if (!sens_io_ok(io, accumulate_only)) return fail(DTMPC_E_BADARG, "missing or misaligned (16 bytes) buffer in sensitivity io");
    SensParams<R> P;
    memset(&P, 0, sizeof(P));
    convert_env(*env, P.env);
    convert_ocp(*ocp, P.ocp);
    convert_sens(*io, P.io);
    P.B = B;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((B + SENS_BLOCK - 1) / SENS_BLOCK);
    if (accumulate_only) {
        ift_accumulate_kernel<R><<<grid, SENS_BLOCK, 0, st>>>(P);
    } else {
        if (!ws) return fail(DTMPC_E_BADARG, "null workspace");
        if (ws_bytes < dtmpc_sens_workspace_bytes(ocp->horizon, B, (int)sizeof(R)))
            return fail(DTMPC_E_WORKSPACE, "sensitivity workspace too small");
        P.ws = (R *)ws;
        launch_sens_kernel<R>(P, grid, st);
    }
    count_launches(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "sensitivity kernel launch");
    return DTMPC_OK;
}

// ---------------------------------------------------------------------------
// deterministic column sums: fixed chunking (1024 rows per block) and fixed trees
// ---------------------------------------------------------------------------
constexpr int RED_BLOCK = 256;
constexpr int RED_ROWS = 1024;

template <typename R> __device__ __forceinline__ R block_tree_sum(R v, R *sh) {
    // warp tree by shuffles, then a tree over the warp partials: the order depends only on the thread index
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    R t = R(0);
    if (warp == 0) {
        t = lane < (RED_BLOCK / 32) ? sh[lane] : R(0);
        for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
    }
    return t;  // valid in thread 0
}

template <typename R>
__global__ void __launch_bounds__(RED_BLOCK) reduce_partial_kernel(const R *__restrict__ g, long long B, int ncols,
                                                                   R *__restrict__ partial) {
    __shared__ R sh[RED_BLOCK / 32];
    const long long base = (long long)blockIdx.x * RED_ROWS;
    for (int c = 0; c < ncols; ++c) {
        R v = R(0);
        for (int j = 0; j < RED_ROWS / RED_BLOCK; ++j) {
            const long long row = base + (long long)j * RED_BLOCK + threadIdx.x;
            if (row < B) v += g[row * ncols + c];
        }
        const R t = block_tree_sum(v, sh);
        if (threadIdx.x == 0) partial[(long long)blockIdx.x * ncols + c] = t;
    }
}

template <typename R>
__global__ void __launch_bounds__(RED_BLOCK) reduce_final_kernel(const R *__restrict__ partial, long long nblocks,
                                                                 int ncols, R *__restrict__ out) {
    __shared__ R sh[RED_BLOCK / 32];
    for (int c = 0; c < ncols; ++c) {
        R v = R(0);
        for (long long b = threadIdx.x; b < nblocks; b += RED_BLOCK) v += partial[b * ncols + c];
        const R t = block_tree_sum(v, sh);
        if (threadIdx.x == 0) out[c] = t;
    }
}

template <typename R>
int reduce_impl(const void *g, long long B, int ncols, void *out, void *ws, size_t ws_bytes, void *stream) {
    if (!g || !out || ncols < 1 || B < 0) return fail(DTMPC_E_BADARG, "bad reduce arguments");
    if (!ws || ws_bytes < dtmpc_reduce_workspace_bytes(B, ncols, (int)sizeof(R)))
        return fail(DTMPC_E_WORKSPACE, "reduce workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nb = B > 0 ? (B + RED_ROWS - 1) / RED_ROWS : 0;
    if (nb > 0) reduce_partial_kernel<R><<<(unsigned)nb, RED_BLOCK, 0, st>>>((const R *)g, B, ncols, (R *)ws);
    reduce_final_kernel<R><<<1, RED_BLOCK, 0, st>>>((const R *)ws, nb, ncols, (R *)out);
    count_launches(nb > 0 ? 2 : 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "reduce kernel launch");
    return DTMPC_OK;
}

// ---------------------------------------------------------------------------
// rollout, point-wise primitives, closed-loop plant step
// ---------------------------------------------------------------------------
template <typename R>
__global__ void rollout_kernel(const __grid_constant__ EnvT<R> E, const __grid_constant__ OcpT<R> O, long long B,
                               const R *__restrict__ x0, const R *__restrict__ V, const R *__restrict__ theta,
                               long long theta_stride, R *__restrict__ X) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int T = O.T;
    Theta<R> th;
    load_theta(th, theta + i * theta_stride, O);
    const R *xp = x0 + 4 * i;
    State<R> x{xp[0], xp[1], xp[2], xp[3], R(0)};
    x.Bc = barrier_of_state(E, th, x.px, x.py);
    R *Xo = X + i * (4LL * (T + 1));
    Xo[0] = x.px; Xo[1] = x.py; Xo[2] = x.th; Xo[3] = x.b;
    for (int k = 0; k < T; ++k) {
        const R u0 = V[i * (2LL * T) + 2 * k], u1 = V[i * (2LL * T) + 2 * k + 1];
        step(E, th, x, u0, u1);
        Xo[4 * (k + 1)] = x.px; Xo[4 * (k + 1) + 1] = x.py; Xo[4 * (k + 1) + 2] = x.th; Xo[4 * (k + 1) + 3] = x.b;
    }
}

template <typename R>
__global__ void primitives_kernel(const __grid_constant__ EnvT<R> E, const __grid_constant__ OcpT<R> O, long long B,
                                  const R *__restrict__ xh, const R *__restrict__ u, const R *__restrict__ theta,
                                  long long theta_stride, const R *__restrict__ target, R *x_next, R *A, R *Bm, R *h,
                                  R *grad_h, R *cost, R *u_clamped, int *active, R *b0) {
    using m = M<R>;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    Theta<R> th;
    load_theta(th, theta + i * theta_stride, O);
    const R px = xh[4 * i], py = xh[4 * i + 1], pt = xh[4 * i + 2], pb = xh[4 * i + 3];
    const R u0 = u[2 * i], u1 = u[2 * i + 1];
    State<R> x{px, py, pt, pb, R(0)};
    x.Bc = barrier_of_state(E, th, px, py);
    step(E, th, x, u0, u1);
    if (x_next) {
        x_next[4 * i] = x.px; x_next[4 * i + 1] = x.py; x_next[4 * i + 2] = x.th; x_next[4 * i + 3] = x.b;
    }
    R hv, hx, hy;
    eval_h<R, true>(E, px, py, hv, hx, hy);
    if (h) h[i] = hv;
    if (b0) b0[i] = barrier_B(hv - th.tight, th.a_jac, th.eps_jac);
    if (grad_h) {
        grad_h[3 * i] = hx; grad_h[3 * i + 1] = hy; grad_h[3 * i + 2] = R(0);
    }
    if (A || Bm) {
        R gcx, gcy, gnx, gny, t0;
        barrier_grad(E, th, px, py, gcx, gcy, t0);
        barrier_grad(E, th, x.px, x.py, gnx, gny, t0);
        Lin<R> L;
        linearize(E, th, pt, u0, gnx, gny, gcx, gcy, L);
        if (A) {
            R *a = A + 16 * i;
            a[0] = 1; a[1] = 0; a[2] = L.a02; a[3] = 0;
            a[4] = 0; a[5] = 1; a[6] = L.a12; a[7] = 0;
            a[8] = 0; a[9] = 0; a[10] = 1; a[11] = 0;
            a[12] = L.r0; a[13] = L.r1; a[14] = L.r2; a[15] = th.gamma;
        }
        if (Bm) {
            R *b = Bm + 8 * i;
            b[0] = L.b00; b[1] = 0; b[2] = L.b10; b[3] = 0; b[4] = 0; b[5] = E.dt; b[6] = L.s0; b[7] = 0;
        }
    }
    if (cost && target)
        cost[i] = stage_cost(th, px, py, pt, pb, u0, u1, target[0], target[1], target[2], R(0), R(0));
    R c0 = u0, c1 = u1;
    clamp_u(E, c0, c1);
    if (u_clamped) {
        u_clamped[2 * i] = c0; u_clamped[2 * i + 1] = c1;
    }
    if (active) {
        bool a0, a1;
        active_u(E, u0, u1, a0, a1);
        active[2 * i] = a0; active[2 * i + 1] = a1;
    }
    (void)m::abs_(R(0));
}

// Counter-based disturbance generator: Philox4x32-10 (Salmon et al., SC'11) keyed by the seed, counter = (global
// instance index, closed-loop step, draw).  A draw is a pure function of (seed, instance, step): independent of the
// batch size, of the sharding across ranks and of CUDA-graph replay (the step counter lives in device memory).
struct Philox4 { unsigned v[4]; };
__host__ __device__ inline Philox4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}
template <typename R> struct RngT {
    R low[3], high[3];
    unsigned long long seed;
    long long first;          // global index of instance 0
    const long long *step;    // device counter t
    int enabled;
};
__device__ inline double unit_from_bits(unsigned lo, unsigned hi, double) {   // 53 bits, [0, 1)
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * 0x1.0p-53;
}
__device__ inline float unit_from_bits(unsigned lo, unsigned, float) { return (float)(lo >> 8) * 0x1.0p-24f; }   // 24 bits

// reference core/tube_mpc.py:990-1023; disturbance core/systems/dubins.py:57-65 (w = low + (high - low) * U[0,1)^3)
template <typename R>
__global__ void plant_step_kernel(const __grid_constant__ EnvT<R> E, const __grid_constant__ OcpT<R> On,
                                  const __grid_constant__ OcpT<R> Oa, long long B, const R *__restrict__ Un,
                                  const R *__restrict__ Ua, const R *__restrict__ theta_n, long long sn,
                                  const R *__restrict__ theta_a, long long sa, const R *__restrict__ w,
                                  const __grid_constant__ RngT<R> G, R *xh, R *xhb, R *Vn_ws, R *Va_ws) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int T = On.T;
    Theta<R> tn, ta;
    load_theta(tn, theta_n + i * sn, On);
    load_theta(ta, theta_a + i * sa, Oa);
    // the closed loop propagates b with the gradient-time DBaS config (YAML eps): tube_mpc.py:609, 614
    tn.eps_dyn = tn.eps_jac; tn.a_dyn = tn.a_jac;
    ta.eps_dyn = ta.eps_jac; ta.a_dyn = ta.a_jac;
    {
        State<R> x{xh[4 * i], xh[4 * i + 1], xh[4 * i + 2], xh[4 * i + 3], R(0)};
        x.Bc = barrier_of_state(E, ta, x.px, x.py);
        step(E, ta, x, Ua[i * (2LL * T)], Ua[i * (2LL * T) + 1]);
        R w0 = R(0), w1 = R(0), w2 = R(0);
        if (w) {
            w0 = w[3 * i]; w1 = w[3 * i + 1]; w2 = w[3 * i + 2];
        } else if (G.enabled) {
            const unsigned long long gi = (unsigned long long)(G.first + i), t = (unsigned long long)*G.step;
            const unsigned k0 = (unsigned)G.seed, k1 = (unsigned)(G.seed >> 32);
            const Philox4 a = philox4x32_10((unsigned)gi, (unsigned)(gi >> 32), (unsigned)t, 0u, k0, k1);
            const Philox4 b = philox4x32_10((unsigned)gi, (unsigned)(gi >> 32), (unsigned)t, 1u, k0, k1);
            w0 = G.low[0] + (G.high[0] - G.low[0]) * unit_from_bits(a.v[0], a.v[1], R(0));
            w1 = G.low[1] + (G.high[1] - G.low[1]) * unit_from_bits(a.v[2], a.v[3], R(0));
            w2 = G.low[2] + (G.high[2] - G.low[2]) * unit_from_bits(b.v[0], b.v[1], R(0));
        }
        xh[4 * i] = x.px + w0; xh[4 * i + 1] = x.py + w1; xh[4 * i + 2] = x.th + w2; xh[4 * i + 3] = x.b;
    }
    {
        State<R> x{xhb[4 * i], xhb[4 * i + 1], xhb[4 * i + 2], xhb[4 * i + 3], R(0)};
        x.Bc = barrier_of_state(E, tn, x.px, x.py);
        step(E, tn, x, Un[i * (2LL * T)], Un[i * (2LL * T) + 1]);
        xhb[4 * i] = x.px; xhb[4 * i + 1] = x.py; xhb[4 * i + 2] = x.th; xhb[4 * i + 3] = x.b;
    }
    // warm-start shift cat(U[1:], U[-1:])  (tube_mpc.py:1016-1020)
    for (int k = 0; k < T; ++k) {
        const int s = k + 1 < T ? k + 1 : T - 1;
        Vn_ws[i * (2LL * T) + 2 * k] = Un[i * (2LL * T) + 2 * s];
        Vn_ws[i * (2LL * T) + 2 * k + 1] = Un[i * (2LL * T) + 2 * s + 1];
        Va_ws[i * (2LL * T) + 2 * k] = Ua[i * (2LL * T) + 2 * s];
        Va_ws[i * (2LL * T) + 2 * k + 1] = Ua[i * (2LL * T) + 2 * s + 1];
    }
}

__global__ void math_selftest_kernel(const double *__restrict__ x, long long n, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using m = M<double>;
    const double v = x[i], a = ::fabs(v);
    double s, c;
    m::sincos_(v, &s, &c);
    double *o = out + 6 * i;
    o[0] = m::exp_nonpos_(-a);
    o[1] = m::log_ge1_(1.0 + a);
    o[2] = m::log_1to2_(1.0 + (a - ::floor(a)));
    o[3] = s;
    o[4] = c;
    o[5] = m::rcp_(v);
}

template <typename R> __global__ void fma_peak_kernel(R *out, int iters) {
    R a0 = R(1.0) + R(threadIdx.x) * R(1e-9), a1 = a0 + R(1), a2 = a0 + R(2), a3 = a0 + R(3);
    R a4 = a0 + R(4), a5 = a0 + R(5), a6 = a0 + R(6), a7 = a0 + R(7);
    const R b = R(0.999999), c = R(1e-7);
    for (int i = 0; i < iters; ++i) {
        a0 = M<R>::fma_(a0, b, c); a1 = M<R>::fma_(a1, b, c); a2 = M<R>::fma_(a2, b, c); a3 = M<R>::fma_(a3, b, c);
        a4 = M<R>::fma_(a4, b, c); a5 = M<R>::fma_(a5, b, c); a6 = M<R>::fma_(a6, b, c); a7 = M<R>::fma_(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

template <typename R> double peak_impl(int iters, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 256, grid = sm_count() * 8;
    R *buf = nullptr;
    if (cudaMalloc(&buf, sizeof(R) * block * grid) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    count_launches(1);
    fma_peak_kernel<R><<<grid, block, 0, st>>>(buf, iters);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        fma_peak_kernel<R><<<grid, block, 0, st>>>(buf, iters);
        count_launches(1);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * (double)iters * block * grid / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

}  // namespace

// ---------------------------------------------------------------------------
extern "C" {

int dtmpc_abi_version(void) { return DTMPC_ABI_VERSION; }
const char *dtmpc_last_error(void) { return g_err; }

size_t dtmpc_solve_workspace_bytes(int32_t T, int64_t B, int32_t n_alpha, int32_t elem_size) {
    if (T < 1 || B < 0 || n_alpha < 1 || n_alpha > DTMPC_MAX_ALPHAS) return 0;
    if (elem_size == 8) return carve_solve_ws<double>(nullptr, T, B, n_alpha, nullptr) + chunked_ws_bytes_slack();
    if (elem_size == 4) return carve_solve_ws<float>(nullptr, T, B, n_alpha, nullptr) + chunked_ws_bytes_slack();
    return 0;
}
size_t dtmpc_step_workspace_bytes(int32_t T, int64_t B, int32_t n_alpha, int32_t elem_size) {
    const size_t a = dtmpc_solve_workspace_bytes(T, B, n_alpha, elem_size);
    if (a == 0) return 0;
    // per chunk: sens_ws_rows(T) rows of whole tiles, aligned; chunks add at most one tile and one alignment each
    return a + (size_t)sens_ws_rows(T) * (size_t)((B + 31) / 32 * 32 + 32 * MAX_CHUNKS) * (size_t)elem_size +
           256 * (size_t)MAX_CHUNKS;
}
uint64_t dtmpc_kernel_launches(void) { return g_kernel_launches.load(std::memory_order_relaxed); }
size_t dtmpc_sens_workspace_bytes(int32_t T, int64_t B, int32_t elem_size) {
    if (T < 1 || B < 0) return 0;
    return (size_t)sens_ws_rows(T) * (size_t)((B + 31) / 32 * 32) * (size_t)elem_size + 256;  // whole tiles
}
size_t dtmpc_reduce_workspace_bytes(int64_t B, int32_t ncols, int32_t elem_size) {
    const long long nb = B > 0 ? (B + RED_ROWS - 1) / RED_ROWS : 0;
    return (size_t)(nb + 1) * (size_t)ncols * (size_t)elem_size;
}

#define DTMPC_DEFINE(SUFFIX, R)                                                                                       \
    int dtmpc_ilqr_solve_##SUFFIX(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_solve_io *io,    \
                                  void *ws, size_t ws_bytes, void *stream) {                                          \
        return solve_impl<R>(env, ocp, nullptr, B, io, nullptr, ws, ws_bytes, stream);                                \
    }                                                                                                                 \
    int dtmpc_tube_solve_##SUFFIX(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,        \
                                  const dtmpc_solve_io *nio, const dtmpc_solve_io *aio, void *ws, size_t ws_bytes,    \
                                  void *stream) {                                                                     \
        if (!aux || !aio) return fail(DTMPC_E_BADARG, "tube solve needs the auxiliary OCP");                          \
        return solve_impl<R>(env, nom, aux, B, nio, aio, ws, ws_bytes, stream);                                       \
    }                                                                                                                 \
    int dtmpc_tube_step_##SUFFIX(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,         \
                                 const dtmpc_solve_io *nio, const dtmpc_solve_io *aio, const dtmpc_sens_io *sa,       \
                                 const dtmpc_sens_io *sn, void *ws, size_t ws_bytes, void *stream) {                  \
        if (!aux || !aio || !sa) return fail(DTMPC_E_BADARG, "tube step needs the auxiliary OCP and sens_aux");       \
        return solve_impl<R>(env, nom, aux, B, nio, aio, ws, ws_bytes, stream, sa, sn, true);                         \
    }                                                                                                                 \
    int dtmpc_sensitivity_##SUFFIX(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_sens_io *io,    \
                                   void *ws, size_t ws_bytes, void *stream) {                                         \
        return sens_impl<R>(env, ocp, B, io, ws, ws_bytes, stream, false);                                            \
    }                                                                                                                 \
    int dtmpc_ift_accumulate_##SUFFIX(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const dtmpc_sens_io *io, \
                                      void *stream) {                                                                 \
        return sens_impl<R>(env, ocp, B, io, nullptr, 0, stream, true);                                               \
    }                                                                                                                 \
    int dtmpc_reduce_columns_##SUFFIX(const void *g, int64_t B, int32_t ncols, void *out, void *ws, size_t ws_bytes,  \
                                      void *stream) {                                                                 \
        return reduce_impl<R>(g, B, ncols, out, ws, ws_bytes, stream);                                                \
    }                                                                                                                 \
    int dtmpc_rollout_##SUFFIX(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const void *x0, const void *V,  \
                               const void *theta, int64_t theta_stride, void *X, void *stream) {                      \
        if (!env_ok(env) || !ocp_ok(ocp) || B < 0) return fail(DTMPC_E_BADARG, "bad env/ocp description");            \
        if (B == 0) return DTMPC_OK;                                                                                  \
        if (!x0 || !V || !theta || !X) return fail(DTMPC_E_BADARG, "null buffer in rollout");                         \
        EnvT<R> E;                                                                                                    \
        OcpT<R> O;                                                                                                    \
        convert_env(*env, E);                                                                                         \
        convert_ocp(*ocp, O);                                                                                         \
        count_launches(1);                                                                                            \
        rollout_kernel<R><<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(                             \
            E, O, B, (const R *)x0, (const R *)V, (const R *)theta, theta_stride, (R *)X);                            \
        cudaError_t e = cudaGetLastError();                                                                           \
        return e == cudaSuccess ? DTMPC_OK : cuda_fail(e, "rollout kernel launch");                                   \
    }                                                                                                                 \
    int dtmpc_primitives_##SUFFIX(const dtmpc_env *env, const dtmpc_ocp *ocp, int64_t B, const void *xh,              \
                                  const void *u, const void *theta, int64_t theta_stride, const void *target,         \
                                  void *x_next, void *A, void *Bm, void *h, void *grad_h, void *cost,                 \
                                  void *u_clamped, int32_t *active, void *b0, void *stream) {                         \
        if (!env_ok(env) || !ocp_ok(ocp) || B < 0) return fail(DTMPC_E_BADARG, "bad env/ocp description");            \
        if (B == 0) return DTMPC_OK;                                                                                  \
        if (!xh || !u || !theta) return fail(DTMPC_E_BADARG, "null buffer in primitives");                            \
        EnvT<R> E;                                                                                                    \
        OcpT<R> O;                                                                                                    \
        convert_env(*env, E);                                                                                         \
        convert_ocp(*ocp, O);                                                                                         \
        count_launches(1);                                                                                            \
        primitives_kernel<R><<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(                          \
            E, O, B, (const R *)xh, (const R *)u, (const R *)theta, theta_stride, (const R *)target, (R *)x_next,     \
            (R *)A, (R *)Bm, (R *)h, (R *)grad_h, (R *)cost, (R *)u_clamped, active, (R *)b0);                       \
        cudaError_t e = cudaGetLastError();                                                                           \
        return e == cudaSuccess ? DTMPC_OK : cuda_fail(e, "primitives kernel launch");                                \
    }                                                                                                                 \
    int dtmpc_plant_step_##SUFFIX(const dtmpc_env *env, const dtmpc_ocp *nom, const dtmpc_ocp *aux, int64_t B,        \
                                  const void *Un, const void *Ua, const void *theta_n, int64_t sn,                    \
                                  const void *theta_a, int64_t sa, const void *w, const dtmpc_disturbance *rng,       \
                                  void *xh, void *xhb, void *Vn_ws, void *Va_ws, void *stream) {                      \
        if (!env_ok(env) || !ocp_ok(nom) || !ocp_ok(aux) || B < 0)                                                    \
            return fail(DTMPC_E_BADARG, "bad env/ocp description");                                                   \
        if (B == 0) return DTMPC_OK;                                                                                  \
        if (!Un || !Ua || !theta_n || !theta_a || !xh || !xhb || !Vn_ws || !Va_ws)                                    \
            return fail(DTMPC_E_BADARG, "null buffer in plant step");                                                 \
        EnvT<R> E;                                                                                                    \
        OcpT<R> On, Oa;                                                                                               \
        convert_env(*env, E);                                                                                         \
        convert_ocp(*nom, On);                                                                                        \
        convert_ocp(*aux, Oa);                                                                                        \
        RngT<R> G;                                                                                                    \
        memset(&G, 0, sizeof(G));                                                                                     \
        if (!w && rng) {                                                                                              \
            if (!rng->step) return fail(DTMPC_E_BADARG, "disturbance generator without a step counter");              \
            for (int c = 0; c < 3; ++c) {                                                                             \
                G.low[c] = (R)rng->low[c];                                                                            \
                G.high[c] = (R)rng->high[c];                                                                          \
            }                                                                                                         \
            G.seed = rng->seed;                                                                                       \
            G.first = rng->first_instance;                                                                            \
            G.step = (const long long *)rng->step;                                                                    \
            G.enabled = 1;                                                                                            \
        }                                                                                                             \
        count_launches(1);                                                                                            \
        plant_step_kernel<R><<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(                          \
            E, On, Oa, B, (const R *)Un, (const R *)Ua, (const R *)theta_n, sn, (const R *)theta_a, sa,               \
            (const R *)w, G, (R *)xh, (R *)xhb, (R *)Vn_ws, (R *)Va_ws);                                              \
        cudaError_t e = cudaGetLastError();                                                                           \
        return e == cudaSuccess ? DTMPC_OK : cuda_fail(e, "plant step kernel launch");                                \
    }

DTMPC_DEFINE(f64, double)
DTMPC_DEFINE(f32, float)

int dtmpc_math_selftest_f64(const void *x, int64_t n, void *out, void *stream) {
    if (n < 0 || (n > 0 && (!x || !out))) return fail(DTMPC_E_BADARG, "bad selftest arguments");
    if (n == 0) return DTMPC_OK;
    count_launches(1);
    math_selftest_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const double *)x, n, (double *)out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DTMPC_OK : cuda_fail(e, "selftest kernel launch");
}

double dtmpc_measure_fp64_peak_tflops(int32_t iters, void *stream) { return peak_impl<double>(iters, stream); }
double dtmpc_measure_fp32_peak_tflops(int32_t iters, void *stream) { return peak_impl<float>(iters, stream); }

}  // extern "C"
