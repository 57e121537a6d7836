// foto_kernels.cuh -- launch wrappers of the sm_100a kernels (definitions in *.cu).
#pragma once
#include "common.cuh"

namespace foto {

struct Dims {
    int Nt, Ny, Nx;     // Nt = time planes held locally
    unsigned int P;     // Nx*Ny
    unsigned int N;     // Nt*P   (grids up to 2^31-1 cells)
    // time-slab view (foto_slab_*): this rank owns the global planes [n0, n0 + Nt) of gNt; 3-component
    // fields have component stride cs, and planes -1 and Nt (halos) are addressable when they exist
    // globally.  Whole-volume calls use n0 = 0, gNt = Nt, cs = N.
    int n0, gNt;
    unsigned int cs;
    // outer-loop pipelining (foto_solve_dev): kernels of a speculatively enqueued ALG2 iteration return at once when
    // *skip != 0 (the iteration before met the stopping rule).  NULL: always run.
    const int *skip = nullptr;
};

// device-side record of the ALG2 outer loop (benamou_brenier.py:204-258), written by k_outer_decide
struct OuterState {
    double crit_prev;       // criterion of the previous outer iteration
    int done;               // stopping rule met (or max_it reached)
    int n_outer;            // outer iterations completed
};
struct OuterTrace { double *crit; int *cg_iters; int *cg_info; };   // device arrays, one entry per outer iteration

// ---- FOTO pointwise / stencil kernels (foto_kernels.cu) --------------------------------
// mu <- [linear-in-time density | 0 | 0],  q <- 0          (benamou_brenier.py:191-194)
void launch_init_state(cudaStream_t st, Dims d, const double *rho0, const double *rhoT, double *mu, double *q);
// K1: F = div_st(mu - r q) + time-boundary terms            (benamou_brenier.py:64-82)
void launch_rhs(cudaStream_t st, Dims d, const double *mu, const double *q, const double *rho0,
                const double *rhoT, double r, double *F);
// K3: grad_st phi, stepB, stepC, clamp, criterion partial sums (benamou_brenier.py:213-251)
// partials: 2*blocks doubles; returns the number of blocks used.
// Returns the number of blocks used (= partial-sum pairs written), or -1 on a launch error (foto_last_error).
int launch_prox_dual(cudaStream_t st, Dims d, const double *phi, double *mu, double *q, double r,
                     double *partials, int max_blocks, int num_sms, int *variant = nullptr);   // variant: 0 register-marching, 1 TMA
// TMA-staged K3 (prox_tma.cu)
bool prox_tma_eligible(const Dims &d, const double *phi, const double *mu, const double *q);
int launch_prox_dual_tma(cudaStream_t st, Dims d, const double *phi, double *mu, double *q, double r, double *partials,
                         int max_blocks, int num_sms, int *blocks_out);
void launch_crit_final(cudaStream_t st, const double *partials, int blocks, double *out2);
// criterion of outer iteration `it` from the K3 partial sums + the reference's stopping rule
// (benamou_brenier.py:246-258), decided on the device; cg_out = the (iterations, info) pair the Poisson kernel wrote
void launch_outer_decide(cudaStream_t st, const double *partials, int blocks, double *crit_sums2, const int *cg_out,
                         OuterState *state, OuterTrace trace, int it, double tol, int max_it);
// stepB alone (benamou_brenier.py:93-149)
void launch_stepB(cudaStream_t st, unsigned int N, const double *p, double *q);
// K4: trajectories + luminosity (utils.py:44-99,148-183)
void launch_flow(cudaStream_t st, Dims d, const double *phi, double *u, double *v, double *m);
// K7: utils.apply_opticalflow (utils.py:186-248); g = scratch P doubles
void launch_warp(cudaStream_t st, int w, int h, const double *f1, const double *u, const double *v,
                 const double *m_or_null, double *g, double *out);
// next-tier rows: .flo payload packing, EE/AE metric sums
void launch_pack_flo(cudaStream_t st, unsigned int n, const double *u, const double *v, float *out);
void launch_flow_metrics(cudaStream_t st, unsigned int n, const double *u, const double *v, const double *ug,
                         const double *vg, double *partials, double *out6);
void launch_ingest_u8(cudaStream_t st, unsigned int n, const unsigned char *in, double *out);
// time-slab transpose: [L][Ny][Nx] planes <-> all-to-all buffer [rank g][L][rows of g][Nx] (rows split like slab.split)
void launch_slab_pack(cudaStream_t st, int L, int Ny, int Nx, int world, const double *planes_in, double *buf_out,
                      double *planes_out, const double *buf_in);
// sum (255 a - 255 b)^2 (utils.IE, utils.py:354); partials: >= 1184 doubles
void launch_ie_sumsq(cudaStream_t st, unsigned int n, const double *a, const double *b, double *partials, double *out1);
// generic tridiagonal-along-one-axis apply used by foto_op_apply
void launch_axis_apply(cudaStream_t st, const double *in, double *out, const double *lo, const double *di,
                       const double *up, int transpose, unsigned int stride, int len, unsigned int total,
                       int accumulate);

// ---- persistent CG for stepA (cg_kernels.cu) -------------------------------------------
struct CgArgs {
    const double *b;        // right-hand side F (N)
    double *x;              // solution phi (N)
    double *r, *p0, *p1, *q;// work vectors (N each)
    int Nt, Ny, Nx;
    double rcoef;           // r
    double eps;             // reg_epsilon
    double rtol;
    int maxiter;
    SyncState sync;
    int *out;               // [0] iterations, [1] info (0 converged / maxiter)
    const int *skip = nullptr;   // see Dims::skip
};
// Returns FOTO_OK or an error code.  grid/block are chosen by cg_stream_config().
int cg_stream_config(int device, int *grid, int *block);
int launch_cg_stream(cudaStream_t st, const CgArgs &a, int grid, int block);

// scratch of the on-chip resident kernels (state in shared memory/registers, one tile per SM), owned by the context
struct OnchipScratch {
    int num_sms = 0;
    size_t smem_optin = 0;
    long long *prof = nullptr;  // 8 cycle counters per CTA (debugging aid, see foto_debug_onchip_prof)
    double *fused_edges = nullptr; size_t fused_edges_bytes = 0;   // cg_fused.cu
    unsigned long long *fused_slots = nullptr, *fused_slots_raw = nullptr;
    unsigned int fused_launch_seq = 0;
    bool fused_attr_set = false;
    double *gnf_edges = nullptr; size_t gnf_edges_bytes = 0;    // gn_fused.cu
    unsigned long long *gnf_slots = nullptr;
    bool gnf_attr_set = false;
};
void onchip_release(OnchipScratch &s);
// on-chip resident, one grid all-reduce per iteration (Chronopoulos-Gear arrangement); cg_fused.cu
bool cg_fused_fits(OnchipScratch &s, int device, int Nt, int Ny, int Nx);
int launch_cg_fused(cudaStream_t st, const CgArgs &a, int device, OnchipScratch &s);

// ---- exact Poisson solve by separable DCT (dct_kernels.cu) -------------------------------
struct DctTables {              // device pointers, owned by the context, valid for (Nt, Ny, Nx)
    int Nt = 0, Ny = 0, Nx = 0;
    double *base = nullptr;     // one allocation holding everything below
    double *Cx = nullptr, *CxT = nullptr, *Cy = nullptr, *CyT = nullptr, *Ct = nullptr, *CtT = nullptr;
    double *lam_x = nullptr, *lam_y = nullptr, *lam_t = nullptr;
    // even / odd folded transforms (Nx and Ny multiples of 4): Ex[b][j][i] = Cx[2j+b][i] for i < Nx/2, ExT its transposes,
    // the same for y, and the eigenvalues in the permuted spectrum order (even frequencies, then odd)
    bool split = false;
    double *Ex = nullptr, *ExT = nullptr, *Ey = nullptr, *EyT = nullptr, *lam_xp = nullptr, *lam_yp = nullptr;
    // second folding level of an axis whose length is a multiple of 8 (lx / ly = 2): E2x[c][m][i] = Cx[4m+2c][i], i < Nx/4
    int lx = 1, ly = 1;
    double *E2x = nullptr, *E2xT = nullptr, *E2y = nullptr, *E2yT = nullptr;
};
void dct_host_folded2(int n, const std::vector<double> &C, const std::vector<double> &lam, std::vector<double> &E2,
                      std::vector<double> &E2T, std::vector<double> &lam_p2);
void dct_host_folded(int n, const std::vector<double> &C, const std::vector<double> &lam, std::vector<double> &E,
                     std::vector<double> &ET, std::vector<double> &lam_p);
void dct_host_tables(int n, std::vector<double> &C, std::vector<double> &Ct, std::vector<double> &lam);
int launch_poisson_dct(cudaStream_t st, const DctTables &tb, int Nt, int Ny, int Nx, double r, double eps,
                       const double *F, double *phi, double *w0, double *w1);
int launch_dct_xy(cudaStream_t st, const DctTables &tb, int nplanes, int Ny, int Nx, const double *in, double *out,
                  double *tmp, int inverse);
int launch_dct_t_solve(cudaStream_t st, const DctTables &tb, int Nt, int ny_loc, int Nx, int y_off, double r, double eps,
                       const double *in, double *out);

// ---- truncated CG over time slabs, stepwise (cg_slab.cu): the collectives between the steps belong to the caller
size_t cg_slab_state_words();
size_t cg_slab_partial_words();
int launch_cg_slab(cudaStream_t st, int op, int gNt, int n0, int nloc, int Ny, int Nx, double rcoef, double eps, double rtol, int it,
                   int maxiter, const double *b, double *x, double *r, double *p_old, double *p_new, double *q, double *partials,
                   double *state);

// ---- Gennert-Negahdaripour (gn_kernels.cu) ---------------------------------------------
// K5: fx, fy (central, zero on the border), ft, Jacobi inverse diagonal, right-hand side
void launch_gn_coeffs(cudaStream_t st, int w, int h, const double *f1, const double *f2, double alpha,
                      double lam, double *fx, double *fy, double *dinv, double *b);
void launch_gn_apply(cudaStream_t st, int w, int h, const double *fx, const double *fy, const double *f2,
                     double alpha, double lam, const double *x, double *y);
struct GnArgs {
    const double *fx, *fy, *f2, *dinv, *b;   // P, P, P, 3P, 3P
    double *x, *r, *z, *p0, *p1, *q;         // 3P each
    int w, h;
    double alpha, lam, rtol;
    int maxiter;
    SyncState sync;
    int *out;                                // [0] iterations, [1] info
};
int gn_pcg_config(int device, int *grid, int *block);
int launch_gn_pcg(cudaStream_t st, const GnArgs &a, int grid, int block);
// spectral (DCT) preconditioner applied in TF32 on the tensor cores, fp64 Krylov recurrences (gn_dct.cu)
struct GnDctTables {            // fp32 zero-padded DCT matrices, owned by the context, valid for (w, h)
    int w = 0, h = 0, wp = 0, hp = 0;
    float *base = nullptr, *Cx = nullptr, *CxT = nullptr, *Cy = nullptr, *CyT = nullptr;
    // even / odd folded transforms (w and h even): half sizes padded to multiples of 4 and the folded matrices
    // Ex[b][j][i] = Cx[2j+b][i] (i, j < w/2), ExT its transposes, the same for y
    bool fold = false;
    int wq = 0, hq = 0;
    float *Ex = nullptr, *ExT = nullptr, *Ey = nullptr, *EyT = nullptr;
    size_t volume_floats() const            // one fp32 work volume: natural [3][hp][wp] or folded [3][2 hq][2 wq] layouts
    {
        const size_t rows = fold && 2 * hq > hp ? 2 * hq : hp, cols = fold && 2 * wq > wp ? 2 * wq : wp;
        return 3 * rows * cols;
    }
};
struct GnDctArgs {
    const double *fx, *fy, *f2, *b;         // P, P, P, 3P
    const double *lam_x, *lam_y;            // eigenvalues of -lap1d (DctTables)
    const GnDctTables *tb;
    double *x, *r, *p, *s, *wv;             // 3P each
    double *gbar, *partials6, *partials3;   // 8, 6 * 1184, 3 * 592
    float *r32, *t1, *t2, *u32;             // tb->volume_floats() each
    void *state;                            // gn_dct_state_bytes()
    int w, h, maxiter;
    double alpha, lam, rtol;
};
size_t gn_dct_state_bytes();
int gn_dct_prepare_tables(cudaStream_t st, const DctTables &tb, int w, int h, GnDctTables &out);
int gn_dct_begin(cudaStream_t st, const GnDctArgs &a);
int gn_dct_enqueue_iterations(cudaStream_t st, const GnDctArgs &a, int it0, int count, int *launches);
void gn_dct_copy_out(cudaStream_t st, const GnDctArgs &a, double *u, double *v, double *m);
void gn_dct_read_state(const void *host_copy, int *done, int *iters, int *info);
// on-chip resident, one grid all-reduce per iteration (gn_fused.cu): uses fx, fy, f2, dinv, b, x, out, sync.error only
bool gn_fused_fits(OnchipScratch &s, int device, int h, int w);
int launch_gn_fused(cudaStream_t st, const GnArgs &a, int device, OnchipScratch &s);

}  // namespace foto
