/*
 * foto_b200.h -- C ABI of libfoto_b200.so: the two dense-grid optical-flow solvers of
 * thomasjacumin/optical-flow-optimal-transport as hand-written sm_100a CUDA.
 *
 * The reference is pure Python and has no FFI of its own; each entry point below replaces the
 * reference function named beside it (file:line in the reference repository) and is what a
 * ctypes binding in the reference's modules would call (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - float64 everywhere, C-contiguous.  Flat index k = n*P + y*Nx + x (x fastest), P = Nx*Ny,
 *     N = Nt*P; 3-component fields are concatenated [c0 | c1 | c2] (reference layout,
 *     benamou_brenier.py:119-121,191-192).  Note Nx = image width, Ny = image height.
 *   - "host" entry points take host pointers, are blocking, and own all device traffic
 *     (H2D, kernels, D2H).  "_dev" entry points take device pointers of the context's device
 *     and are blocking too (the outer loop's stopping rule is evaluated on the host).
 *   - Every function returns FOTO_OK (0) or a negative FOTO_ERR_* code; foto_last_error()
 *     gives a thread-local message.  There is NO CPU fallback: without a usable CUDA device
 *     every compute entry point fails with FOTO_ERR_NODEV / FOTO_ERR_CUDA.
 *   - Thread safety: a foto_ctx must not be used from two threads at once; different contexts
 *     (same or different devices) are independent.  The host entry points use one lazily
 *     created context per (thread, current CUDA device).
 */
#ifndef FOTO_B200_H
#define FOTO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FOTO_OK                 0
#define FOTO_ERR_ARG           -1   /* bad argument (sizes < 2, NULL, unknown id)              */
#define FOTO_ERR_CUDA          -2   /* CUDA runtime error, message in foto_last_error()        */
#define FOTO_ERR_NODEV         -3   /* no CUDA device / device is not sm_100                   */
#define FOTO_ERR_BREAKDOWN     -4   /* CG breakdown (reference: RuntimeError, bb.py:88-89)     */
#define FOTO_ERR_TIMEOUT       -5   /* grid-barrier watchdog fired inside a persistent kernel  */
#define FOTO_ERR_NOTIMPL       -6   /* unknown boundary condition / operator (NotImplementedError) */

/* Poisson back-ends for stepA (parity trap #1) */
#define FOTO_POISSON_CG_PARITY  0   /* scipy.sparse.linalg.cg recurrence, rtol 1e-6, maxiter 1000, x0 = 0 */
#define FOTO_POISSON_CG_TIGHT   1   /* same recurrence, rtol 1e-13, maxiter 100000 ("tight" oracle)      */
#define FOTO_POISSON_DCT_EXACT  2   /* exact solve by separable DCT-II/III (dense fp64 transforms); agrees
                                       with CG_TIGHT to ~1e-12, NOT with the reference's truncated CG     */

/* operator ids for foto_op_apply (reference operators.py) */
#define FOTO_OP_GRAD_ST         0   /* operators.py:114-127   N  -> 3N */
#define FOTO_OP_DIV_ST          1   /* operators.py:129-142   3N -> N  */
#define FOTO_OP_LAPLACIAN_ST    2   /* operators.py:144-157   N  -> N  */
#define FOTO_OP_GRAD            3   /* operators.py:160-169   P  -> 2P */
#define FOTO_OP_DIV             4   /* operators.py:182-191   2P -> P  */
#define FOTO_OP_GRAD_FORWARD    5   /* operators.py:171-180   P  -> 2P */
/* 1-D builder ids for foto_tri_coeffs (operators.py:5-110) */
#define FOTO_1D_FORWARD_WEIRD   0
#define FOTO_1D_BACKWARD_WEIRD  1
#define FOTO_1D_CENTRAL_WEIRD   2
#define FOTO_1D_CENTRAL         3
#define FOTO_1D_FORWARD         4
#define FOTO_1D_BACKWARD        5
#define FOTO_1D_LAP             6
#define FOTO_BC_N               0
#define FOTO_BC_D               1

typedef struct foto_ctx foto_ctx;

/* per-context counters, reset by foto_ctx_reset_stats().  Times are CUDA-event milliseconds
 * measured on the context's stream, accumulated only while profiling is on. */
typedef struct foto_stats {
    long long launches;        /* kernels of this library launched                       */
    long long cg_launches;     /* launches of the persistent CG kernel (one per stepA)   */
    long long cg_iterations;   /* CG iterations summed over those launches               */
    long long cg_cells;        /* sum over launches of (iterations * N)                  */
    double    cg_ms;           /* device time inside the CG kernel                       */
    double    rhs_ms;          /* K1                                                     */
    double    prox_ms;         /* K3                                                     */
    double    flow_ms;         /* K4                                                     */
    long long rhs_cells, prox_cells;
    long long gn_launches, gn_iterations, gn_pixels;   /* GN persistent PCG kernel       */
    double    gn_ms;
    int       cg_variant;      /* last Poisson solve: 0 streaming, 2 dct_exact, 3 on-chip single-reduction */
    int       prox_variant;    /* last stepB/stepC launch (K3): 0 register-marching kernel, 1 TMA-staged kernel */
} foto_stats;

const char *foto_last_error(void);
int  foto_version(void);
int  foto_device_count(void);                               /* < 0 on error */

/* ---- contexts (device-resident API) ------------------------------------------------ */
int  foto_ctx_create(int device, foto_ctx **out);
void foto_ctx_destroy(foto_ctx *ctx);
int  foto_ctx_device(const foto_ctx *ctx);
int  foto_ctx_set_profiling(foto_ctx *ctx, int on);         /* CUDA-event timing of K1..K4 */
int  foto_ctx_reset_stats(foto_ctx *ctx);
int  foto_ctx_get_stats(foto_ctx *ctx, foto_stats *out);
int  foto_ctx_set_cg_variant(foto_ctx *ctx, int variant);   /* -1 auto, 0 streaming (textbook CG), 2 on-chip single-reduction CG */
/* Variant used by the contexts behind the host-buffer API (default -1 = auto; the environment
 * variable FOTO_CG_VARIANT sets the initial value). */
int  foto_set_default_cg_variant(int variant);
/* Debugging aid: per-phase cycle counters of the on-chip CG kernel (see api.cu). */
int  foto_debug_onchip_prof(foto_ctx *ctx, int enable, long long *out_1024x8);
/* CUDA-event stopwatch on the context's stream (the stream every kernel of the context is
 * launched on): which = 0 records "start", 1 records "stop"; elapsed synchronises on "stop". */
int  foto_ctx_event_record(foto_ctx *ctx, int which);
int  foto_ctx_event_elapsed_ms(foto_ctx *ctx, double *ms);

/* benamou_brenier.solve (benamou_brenier.py:151-271) on device buffers.
 * d_rho0, d_rhoT: P doubles; d_u, d_v, d_m: P doubles (outputs).
 * crit_trace[max_it], cg_iters[max_it] (host, may be NULL): per outer iteration the stopping
 * criterion printed at benamou_brenier.py:252 and the inner CG iteration count;
 * cg_info[max_it] (host, may be NULL): scipy's `info` (0 converged, maxiter otherwise -> the
 * WARNING of benamou_brenier.py:86-87).  *n_outer = outer iterations performed. */
int  foto_solve_dev(foto_ctx *ctx, const double *d_rho0, const double *d_rhoT,
                    int Nt, int Nx, int Ny, double r, double tol, double eps, int max_it,
                    int poisson_backend, double *d_u, double *d_v, double *d_m,
                    double *crit_trace, int *n_outer, int *cg_iters, int *cg_info);

/* classical.GLLOpticalFlow.assemble + process (classical.py:68-130) on device buffers.
 * The direct SuperLU solve is replaced by a matrix-free Jacobi-preconditioned CG run until
 * ||r|| <= rtol ||b|| (rtol <= 0 selects 1e-13; max_it <= 0 selects 20000). */
int  foto_gn_solve_dev(foto_ctx *ctx, const double *d_f1, const double *d_f2, int w, int h,
                       double alpha, double lambda, double rtol, int max_it,
                       double *d_u, double *d_v, double *d_m, int *iters, int *info);

/* Host-buffer variants on an explicit context (H2D of the inputs, solve, D2H of the results,
 * all on the context's stream; buffers may be pageable or pinned). */
int  foto_solve_host(foto_ctx *ctx, const double *rho0, const double *rhoT, int Nt, int Nx, int Ny,
                     double r, double tol, double eps, int max_it, int poisson_backend,
                     double *u, double *v, double *m,
                     double *crit_trace, int *n_outer, int *cg_iters, int *cg_info);
int  foto_gn_solve_host(foto_ctx *ctx, const double *f1, const double *f2, int w, int h,
                        double alpha, double lambda, double rtol, int max_it,
                        double *u, double *v, double *m, int *iters, int *info);

/* ---- time-slab building blocks (device pointers) -----------------------------------
 * One huge volume split into contiguous time slabs, one per rank: rank owns the global planes
 * [n0, n0 + nloc) of gNt.  The exchange steps (1-plane halos of mu_rho, q_a and phi, the t <-> y
 * all-to-all of the DCT, the all-reduce of the two criterion sums) are collectives issued by the host
 * driver (foto_b200/slab.py over torch.distributed / NCCL); these are the compute steps in between.
 * Pointers address the first OWNED plane; 3-component fields have component stride cs (doubles);
 * halo planes at -1 and nloc must be addressable where they exist globally. */
/* launch on the caller's stream (NULL = legacy default stream), or back on the context's own stream */
int  foto_ctx_set_stream(foto_ctx *ctx, void *cuda_stream, int use_own_stream);
/* K1 on a slab: solve_benamou_brenier_step's right-hand side, benamou_brenier.py:64-82 */
int  foto_slab_rhs_dev(foto_ctx *ctx, const double *d_mu, const double *d_q, unsigned long long cs,
                       const double *d_rho0, const double *d_rhoT, double r, int gNt, int n0, int nloc,
                       int Nx, int Ny, double *d_F);
/* K3 on a slab: stepB + stepC + criterion sums (benamou_brenier.py:213-251); d_out2 = [num, den] of this slab */
int  foto_slab_prox_dev(foto_ctx *ctx, const double *d_phi, double *d_mu, double *d_q, unsigned long long cs,
                        double r, int gNt, int n0, int nloc, int Nx, int Ny, double *d_out2);
/* K2b pieces: x/y DCT of nplanes planes (inverse != 0: DCT-III); t solve on a [gNt][ny_loc][Nx] block of rows y_off.. */
int  foto_dct_xy_dev(foto_ctx *ctx, const double *d_in, double *d_out, double *d_tmp, int nplanes, int gNt,
                     int Ny, int Nx, int inverse);
int  foto_dct_t_solve_dev(foto_ctx *ctx, const double *d_in, double *d_out, int gNt, int Ny, int Nx, int y_off,
                          int ny_loc, double r, double eps);
/* Time-slab form of the reference's truncated CG (benamou_brenier.py:85), one step per call; the caller issues the
 * collectives between the steps on the same stream (foto_b200/slab.py):
 *   op 0 init (x = 0, r = b, p = 0; partial b.b -> state[0])      then all-reduce state[0]
 *   per iteration it:  exchange the boundary planes of r;  op 3 phase A (stop test and beta from state[0]; p_new = p_old beta + r
 *   incl. the halo planes, q = A p_new, partial p.q -> state[0]);  all-reduce state[0];  op 5 phase B (alpha, x and r update,
 *   partial r.r -> state[0]);  all-reduce state[0]
 *   op 7 after maxiter iterations without convergence.
 * Kernels return at once when the state's done flag is up, so iterations may be enqueued ahead of the flag being read.
 * d_r, d_p_old, d_p_new point to plane 0 of [nloc + 2][Ny*Nx] arrays (planes -1 and nloc are halos); d_x, d_q, d_b: [nloc][Ny*Nx];
 * d_state: foto_slab_cg_state_words() doubles, zeroed before op 0: [4] done, [5] iterations, [6] info (scipy's). */
int  foto_slab_cg_dev(foto_ctx *ctx, int op, int gNt, int n0, int nloc, int Ny, int Nx, double r, double eps, double rtol, int it,
                      int maxiter, const double *d_b, double *d_x, double *d_r, double *d_p_old, double *d_p_new, double *d_q,
                      double *d_state);
int  foto_slab_cg_state_words(void);
/* t-slab <-> y-slab transpose of the DCT all-to-all in one pass: direction 0 packs nloc planes [nloc][Ny][Nx] into the
 * send buffer whose block for rank g ([nloc][rows of g][Nx], rows of g = [g Ny/world, (g+1) Ny/world)) is contiguous;
 * direction 1 unpacks a received buffer of that layout into planes */
int  foto_slab_pack_dev(foto_ctx *ctx, int direction, int nloc, int Ny, int Nx, int world, const double *d_in, double *d_out);
/* K4 on device buffers: utils.opticalflow_from_benamoubrenier, utils.py:148 */
int  foto_flow_dev(foto_ctx *ctx, const double *d_phi, int Nt, int Nx, int Ny, double *d_u, double *d_v, double *d_m);

/* ---- host-buffer API (what the reference's Python modules bind) ------------------- */
/* benamou_brenier.solve, benamou_brenier.py:151 */
int  foto_solve(const double *rho0, const double *rhoT, int Nt, int Nx, int Ny,
                double r, double tol, double eps, int max_it, int poisson_backend,
                double *u, double *v, double *m,
                double *crit_trace, int *n_outer, int *cg_iters, int *cg_info);
/* benamou_brenier.stepB, benamou_brenier.py:93 : p[3N] -> q[3N] */
int  foto_stepB(const double *p, int Nt, int Nx, int Ny, double *q);
/* benamou_brenier.solve_benamou_brenier_step, benamou_brenier.py:26 (matrix A = -r L + r eps I
 * and div_st are implied by r, eps and the grid; dt = dx = dy = 1 as in solve()). */
int  foto_stepA(const double *mu, const double *q, const double *rho0, const double *rhoT,
                double r, double eps, int Nt, int Nx, int Ny, int poisson_backend,
                double *phi, int *cg_iters, int *cg_info);
/* right-hand side of stepA alone (benamou_brenier.py:64-82) */
int  foto_rhs(const double *mu, const double *q, const double *rho0, const double *rhoT,
              double r, int Nt, int Nx, int Ny, double *F);
/* utils.opticalflow_from_benamoubrenier, utils.py:148 with grad(...,'N'), div(...,'D') */
int  foto_flow_from_phi(const double *phi, int Nt, int Nx, int Ny, double *u, double *v, double *m);
/* operators.* applied matrix-free: out = Op(in) or Op^T(in) */
int  foto_op_apply(int op_id, int bc, int Nt, int Nx, int Ny, double dt, double dx, double dy,
                   int transpose, const double *in, double *out);
/* 1-D builders of operators.py:5-110 as (lo, di, up) rows; host-only helper (no GPU) */
int  foto_tri_coeffs(int kind, int n, double h, int bc, double *lo, double *di, double *up);
/* classical.GLLOpticalFlow.assemble().process(), classical.py:68-130 */
int  foto_gn_solve(const double *f1, const double *f2, int w, int h, double alpha, double lambda,
                   double rtol, int max_it, double *u, double *v, double *m, int *iters, int *info);
/* A @ x and b of the GN system (classical.py:106-110), matrix-free; x, y, b: 3P doubles */
int  foto_gn_system(const double *f1, const double *f2, int w, int h, double alpha, double lambda,
                    const double *x, double *y, double *b);
/* utils.apply_opticalflow, utils.py:186 ; m may be NULL */
int  foto_warp_apply(const double *f1, const double *u, const double *v, int w, int h,
                     const double *m_or_null, double *out);

/* Middlebury .flo payload (utils.saveFlo, utils.py:273-292): n = w*h pairs of float32 (u, v),
 * interleaved; the 12-byte header (float32 202021.25, int32 w, int32 h) is written by the caller. */
int  foto_pack_flo(const double *u, const double *v, int n, float *out_2n);
/* utils.EE / utils.AE (utils.py:294-338): out6 = [sum EE, sum EE^2, #EE<=50, sum AE, sum AE^2, #AE not NaN];
 * mean = sum/count, stddev = sqrt(sumsq/count - mean^2). */
int  foto_flow_metrics(const double *u, const double *v, const double *uGT, const double *vGT, int n, double *out6);

/* ---- device-resident ingest / egress: consume the solver's device outputs ------------
 * All pointers are device pointers; work is enqueued on the context's stream (no host synchronisation except
 * where a host result is returned). */
/* utils.openGrayscaleImage's conversion (utils.py:39-42): 8-bit grey -> float64 k/255, bit-identical to numpy */
int  foto_ingest_u8_dev(foto_ctx *ctx, const unsigned char *d_u8, int n, double *d_out);
/* utils.saveFlo payload (utils.py:285-292): d_out_2n = n interleaved float32 (u, v) pairs */
int  foto_pack_flo_dev(foto_ctx *ctx, const double *d_u, const double *d_v, int n, float *d_out_2n);
/* utils.EE / utils.AE sums as foto_flow_metrics; d_out6: 6 doubles on the device */
int  foto_flow_metrics_dev(foto_ctx *ctx, const double *d_u, const double *d_v, const double *d_uGT, const double *d_vGT,
                           int n, double *d_out6);
/* utils.apply_opticalflow (utils.py:186-248) on device buffers; d_m may be NULL; then IE's sum of squares
 * (utils.py:340-354) against d_IGT when d_ie_sumsq != NULL: *d_ie_sumsq = sum (255 out - 255 IGT)^2 */
int  foto_warp_dev(foto_ctx *ctx, const double *d_f1, const double *d_u, const double *d_v, int w, int h,
                   const double *d_m_or_null, double *d_out, const double *d_IGT_or_null, double *d_ie_sumsq_or_null);

/* Batched ingest + solve + .flo egress: n_pairs pairs of 8-bit grey frames (what the image files hold; 8x less
 * H2D traffic than float64) through pinned staging, benamou_brenier.solve on the device, results returned as the
 * float32 .flo payload (n_pairs * 2P floats) and, optionally, m (n_pairs * P doubles, may be NULL).  Sharded over
 * devices like foto_solve_batch. */
int  foto_solve_batch_u8(int n_pairs, const unsigned char *f0s, const unsigned char *f1s, int Nt, int Nx, int Ny,
                         double r, double tol, double eps, int max_it, int poisson_backend,
                         const int *device_ids, int n_dev, float *flo_payloads, double *ms_or_null, int *n_outer);

/* Many independent pairs of one shape, sharded over devices by a work queue (one host thread
 * per device, no collective).  rho0s/rhoTs: n_pairs*P doubles,
 * us/vs/ms: n_pairs*P doubles; n_outer[n_pairs]; device_ids[n_dev] (NULL: devices 0..n_dev-1). */
int  foto_solve_batch(int n_pairs, const double *rho0s, const double *rhoTs, int Nt, int Nx, int Ny,
                      double r, double tol, double eps, int max_it, int poisson_backend,
                      const int *device_ids, int n_dev,
                      double *us, double *vs, double *ms, int *n_outer);
int  foto_gn_solve_batch(int n_pairs, const double *f1s, const double *f2s, int w, int h,
                         double alpha, double lambda, double rtol, int max_it,
                         const int *device_ids, int n_dev,
                         double *us, double *vs, double *ms, int *iters);

#ifdef __cplusplus
}
#endif
#endif /* FOTO_B200_H */
