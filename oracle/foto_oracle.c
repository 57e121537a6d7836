/*
 * foto_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C, single-threaded restatement of the reference's dense-grid optical-flow path
 * (thomasjacumin/optical-flow-optimal-transport).  It exists only so that tests/, the
 * smoke test and bench.py's cpu_baseline / --impl reference legs can check and time the
 * CUDA library against something that does not need a GPU.  The product path
 * (optical-flow-optimal-transport_b200/) never links, imports or calls this file.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
 * golden vectors produced by running the unmodified reference (tests/golden/make_golden.py).
 * The reference itself ships no tests or goldens.
 *
 * Third-party arithmetic restated here:
 *   scipy.sparse.linalg.cg (scipy 1.15.2 pinned, 1.18.1 installed; _isolve/iterative.py):
 *     atol = rtol*||b||; r=b (x0=0); loop: if ||r|| < atol return; rho=r.r; p=r+beta p;
 *     q=Ap; alpha=rho/(p.q); x+=alpha p; r-=alpha q.   -> oracle_cg()
 *   scipy.sparse.linalg.spsolve (SuperLU direct solve, classical.py:126): restated as
 *     "solve A x = b to machine precision" with a Jacobi-preconditioned CG run to a
 *     relative residual of 1e-14 -> oracle_gn_solve(); oracle/gn_direct.py holds the
 *     scipy-spsolve variant used to cross-check it.
 *
 * Layout: flat index k = n*P + y*Nx + x (x fastest), P = Nx*Ny,
 * N = Nt*P; fields are float64, structure-of-arrays: mu=[rho|m1|m2], q=[a|b1|b2],
 * grad_st phi = [d_t|d_x|d_y], each 3N long.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX(n, y, x) ((size_t)(n) * P + (size_t)(y) * Nx + (size_t)(x))

/* ------------------------------------------------------------------ 1-D builders
 * kind: 0 grad_1d_forward_weird  (operators.py:5-17)
 *       1 grad_1d_backward_weird (operators.py:19-31)
 *       2 grad_1d_central_weird  (operators.py:33-48)
 *       3 grad_1d_central        (operators.py:52-65)
 *       4 grad_1d_forward        (operators.py:67-79)
 *       5 grad_1d_backward       (operators.py:81-93)
 *       6 lap1d                  (operators.py:95-110)
 * bc: 0 = 'N', 1 = 'D'.  Row i of the n x n matrix is (lo[i], di[i], up[i]) on columns
 * (i-1, i, i+1).  The reference patches boundary rows AFTER dividing by h, so those
 * entries are not scaled (SURVEY.md parity trap #3); reproduced here.
 */
int oracle_tri_coeffs(int kind, int n, double h, int bc, double *lo, double *di, double *up)
{
    if (n < 2 || (bc != 0 && bc != 1) || kind < 0 || kind > 6) return -1;
    for (int i = 0; i < n; i++) { lo[i] = di[i] = up[i] = 0.0; }
    switch (kind) {
    case 0:
        for (int i = 0; i < n; i++) { di[i] = -1.0 / h; if (i + 1 < n) up[i] = 1.0 / h; }
        di[n - 1] = 1.0; lo[n - 1] = -1.0;
        break;
    case 1:
        for (int i = 0; i < n; i++) { di[i] = 1.0 / h; if (i > 0) lo[i] = -1.0 / h; }
        di[0] = -1.0; up[0] = 1.0;
        break;
    case 2:
        for (int i = 0; i < n; i++) { if (i > 0) lo[i] = -0.5 / h; if (i + 1 < n) up[i] = 0.5 / h; }
        if (bc == 0) { di[0] = -1.0; up[0] = 1.0; di[n - 1] = 1.0; lo[n - 1] = -1.0; }
        break;
    case 3:
        for (int i = 0; i < n; i++) { if (i > 0) lo[i] = -0.5 / h; if (i + 1 < n) up[i] = 0.5 / h; }
        if (bc == 0) { up[0] = 0.0; lo[n - 1] = 0.0; }
        break;
    case 4:
        for (int i = 0; i < n; i++) { di[i] = -1.0 / h; if (i + 1 < n) up[i] = 1.0 / h; }
        if (bc == 0) di[n - 1] = 0.0;
        break;
    case 5:
        for (int i = 0; i < n; i++) { di[i] = 1.0 / h; if (i > 0) lo[i] = -1.0 / h; }
        if (bc == 0) di[0] = 0.0;
        break;
    case 6: {
        double s = 1.0 / (h * h);
        for (int i = 0; i < n; i++) { di[i] = -2.0 * s; if (i > 0) lo[i] = s; if (i + 1 < n) up[i] = s; }
        if (bc == 0) { di[0] = -s; up[0] = s; di[n - 1] = -s; lo[n - 1] = s; }
        break; }
    }
    return 0;
}

/* y (+)= T_axis x  on an (n2, n1, n0) box, T tridiagonal along `axis` (0 = fastest).
 * transpose != 0 applies T^T. */
static void axis_apply(const double *lo, const double *di, const double *up, int transpose,
                       int axis, int n0, int n1, int n2, const double *x, double *y, int accumulate)
{
    int n[3] = { n0, n1, n2 };
    size_t stride[3] = { 1, (size_t)n0, (size_t)n0 * n1 };
    int len = n[axis];
    size_t st = stride[axis];
    for (int k = 0; k < n2; k++)
        for (int j = 0; j < n1; j++)
            for (int i = 0; i < n0; i++) {
                size_t idx = (size_t)k * stride[2] + (size_t)j * stride[1] + i;
                int c = axis == 0 ? i : (axis == 1 ? j : k);
                double s = 0.0;
                if (!transpose) {
                    if (c > 0) s += lo[c] * x[idx - st];
                    s += di[c] * x[idx];
                    if (c + 1 < len) s += up[c] * x[idx + st];
                } else {
                    if (c > 0) s += up[c - 1] * x[idx - st];
                    s += di[c] * x[idx];
                    if (c + 1 < len) s += lo[c + 1] * x[idx + st];
                }
                if (accumulate) y[idx] += s; else y[idx] = s;
            }
}

/* op: 0 grad_st (operators.py:114-127)   N -> 3N      1 div_st (129-142)       3N -> N
 *     2 laplacian_st (144-157)           N -> N       3 grad (160-169)          P -> 2P
 *     4 div (182-191)                    2P -> P      5 grad_forward (171-180)  P -> 2P
 * For the 2-D operators Nt is ignored.  transpose swaps the roles of in/out.  */
int oracle_op_apply(int op, int bc, int Nt, int Nx, int Ny, double dt, double dx, double dy,
                    int transpose, const double *in, double *out)
{
    if (op < 0 || op > 5 || (bc != 0 && bc != 1)) return -1;
    int three_d = op <= 2;
    if (!three_d) Nt = 1;
    int mx = Nx > Ny ? Nx : Ny; if (Nt > mx) mx = Nt;
    double *c = (double *)malloc(sizeof(double) * 9 * (size_t)mx);
    double *lt = c, *dtt = c + mx, *ut = c + 2 * mx, *lx = c + 3 * mx, *dxx = c + 4 * mx, *ux = c + 5 * mx,
           *ly = c + 6 * mx, *dyy = c + 7 * mx, *uy = c + 8 * mx;
    int kind = (op == 0 || op == 1) ? 2 : (op == 2 ? 6 : (op == 5 ? 4 : 3));
    int rc = 0;
    if (three_d) rc |= oracle_tri_coeffs(kind, Nt, dt, bc, lt, dtt, ut);
    rc |= oracle_tri_coeffs(kind, Nx, dx, bc, lx, dxx, ux);
    rc |= oracle_tri_coeffs(kind, Ny, dy, bc, ly, dyy, uy);
    if (rc) { free(c); return -1; }
    size_t P = (size_t)Nx * Ny, N = P * Nt;
    int stack_out = (op == 0 || op == 3 || op == 5);        /* block column: [T;X;Y] */
    if (op == 2) {                                          /* sum of three axis operators */
        axis_apply(lt, dtt, ut, transpose, 2, Nx, Ny, Nt, in, out, 0);
        axis_apply(lx, dxx, ux, transpose, 0, Nx, Ny, Nt, in, out, 1);
        axis_apply(ly, dyy, uy, transpose, 1, Nx, Ny, Nt, in, out, 1);
    } else if (stack_out != (transpose != 0)) {             /* one field in, stacked fields out */
        size_t o = 0;
        if (three_d) { axis_apply(lt, dtt, ut, transpose, 2, Nx, Ny, Nt, in, out, 0); o = N; }
        axis_apply(lx, dxx, ux, transpose, 0, Nx, Ny, Nt, in, out + o, 0);
        axis_apply(ly, dyy, uy, transpose, 1, Nx, Ny, Nt, in, out + o + N, 0);
    } else {                                                /* stacked fields in, one field out */
        size_t o = 0;
        if (three_d) { axis_apply(lt, dtt, ut, transpose, 2, Nx, Ny, Nt, in, out, 0); o = N; }
        axis_apply(lx, dxx, ux, transpose, 0, Nx, Ny, Nt, in + o, out, three_d);
        axis_apply(ly, dyy, uy, transpose, 1, Nx, Ny, Nt, in + o + N, out, 1);
    }
    free(c);
    return 0;
}

/* ------------------------------------------------------------------ unit-spacing stencils
 * The solver only ever uses dt=dx=dy=1, bc='N' (benamou_brenier.py:185-187,197-203).   */
static inline double dw(const double *f, size_t k, size_t st, int i, int n)   /* "weird" central */
{
    if (i == 0) return f[k + st] - f[k];
    if (i == n - 1) return f[k] - f[k - st];
    return 0.5 * f[k + st] - 0.5 * f[k - st];
}

/* running-sum form of dw(): s += row entries in column order */
static inline double dw_acc(double s, const double *f, size_t k, size_t st, int i, int n)
{
    if (i == 0) { s += -1.0 * f[k]; s += 1.0 * f[k + st]; }
    else if (i == n - 1) { s += -1.0 * f[k - st]; s += 1.0 * f[k]; }
    else { s += -0.5 * f[k - st]; s += 0.5 * f[k + st]; }
    return s;
}

static inline double lap1(const double *f, size_t k, size_t st, int i, int n)  /* 1-D Neumann */
{
    if (i == 0) return f[k + st] - f[k];
    if (i == n - 1) return f[k - st] - f[k];
    return f[k - st] - 2.0 * f[k] + f[k + st];
}

/* q = A p with A = -r L_st + r eps I assembled as the reference does (benamou_brenier.py:201-203)
 * and applied in scipy's csr_matvec order: one running sum per row over the row's stored
 * entries in column order (t-1, y-1, x-1, diagonal, x+1, y+1, t+1), products a_ij * p_j with
 * a_ij = -r off the diagonal and (-r * L_ii + r * eps) on it.  The order matters only for
 * rounding, but rounding noise is what seeds the t-antisymmetric modes when the right-hand
 * side is exactly t-symmetric (Nt = 2, or synthetic inputs made of exact zeros/ones). */
static void apply_A(const double *p, double r, double eps, int Nt, int Nx, int Ny, double *q)
{
    size_t P = (size_t)Nx * Ny;
    double off = -r * 1.0, reps = r * eps * 1.0;
    for (int n = 0; n < Nt; n++)
        for (int y = 0; y < Ny; y++)
            for (int x = 0; x < Nx; x++) {
                size_t k = IDX(n, y, x);
                double Lii = ((n == 0 || n == Nt - 1) ? -1.0 : -2.0)
                           + (((x == 0 || x == Nx - 1) ? -1.0 : -2.0) + ((y == 0 || y == Ny - 1) ? -1.0 : -2.0));
                double s = 0.0;
                if (n > 0) s += off * p[k - P];
                if (y > 0) s += off * p[k - Nx];
                if (x > 0) s += off * p[k - 1];
                s += (-r * Lii + reps) * p[k];
                if (x < Nx - 1) s += off * p[k + 1];
                if (y < Ny - 1) s += off * p[k + Nx];
                if (n < Nt - 1) s += off * p[k + P];
                q[k] = s;
            }
}

static double dot(const double *a, const double *b, size_t n)
{
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* scipy.sparse.linalg.cg restated (x0 = 0, no preconditioner).  Returns info like scipy:
 * 0 converged, maxiter if the loop was exhausted.  *iters = iterations performed. */
int oracle_cg(const double *b, double r, double eps, int Nt, int Nx, int Ny,
              double rtol, int maxiter, double *x, int *iters)
{
    size_t N = (size_t)Nt * Nx * Ny;
    double bn = sqrt(dot(b, b, N));
    *iters = 0;
    if (bn == 0.0) { memcpy(x, b, N * sizeof(double)); return 0; }
    double atol = rtol * bn;
    double *res = (double *)malloc(3 * N * sizeof(double));
    double *p = res + N, *q = res + 2 * N;
    memcpy(res, b, N * sizeof(double));
    memset(x, 0, N * sizeof(double));
    double rho_prev = 0.0;
    int info = maxiter;
    for (int it = 0; it < maxiter; it++) {
        double rho = dot(res, res, N);
        if (sqrt(rho) < atol) { info = 0; break; }
        if (it > 0) {
            double beta = rho / rho_prev;
            for (size_t i = 0; i < N; i++) p[i] = p[i] * beta + res[i];
        } else {
            memcpy(p, res, N * sizeof(double));
        }
        apply_A(p, r, eps, Nt, Nx, Ny, q);
        double alpha = rho / dot(p, q, N);
        for (size_t i = 0; i < N; i++) { x[i] += alpha * p[i]; res[i] -= alpha * q[i]; }
        rho_prev = rho;
        *iters = it + 1;
    }
    free(res);
    return info;
}

/* F = div_st (mu - r q) with the time-boundary terms (benamou_brenier.py:64,73-82) */
void oracle_rhs(const double *mu, const double *q, const double *rho0, const double *rhoT,
                double r, int Nt, int Nx, int Ny, double *F)
{
    size_t P = (size_t)Nx * Ny, N = P * Nt;
    double *w = (double *)malloc(3 * N * sizeof(double));
    for (size_t i = 0; i < 3 * N; i++) w[i] = mu[i] - r * q[i];
    /* scipy's coo_matvec adds the stored entries of [Dt | Dx | Dy] one after the other into
     * F[row]: t-block first (columns n-1, n+1), then x, then y -- one running sum. */
    for (int n = 0; n < Nt; n++)
        for (int y = 0; y < Ny; y++)
            for (int x = 0; x < Nx; x++) {
                size_t k = IDX(n, y, x);
                double s = 0.0;
                s = dw_acc(s, w, k, P, n, Nt);
                s = dw_acc(s, w + N, k, 1, x, Nx);
                s = dw_acc(s, w + 2 * N, k, Nx, y, Ny);
                F[k] = s;
            }
    for (size_t i = 0; i < P; i++) {
        double g0 = rho0[i] - mu[i] + r * q[i];
        F[i] -= g0;
        size_t k = (size_t)(Nt - 1) * P + i;
        double gN = rhoT[i] - mu[k] + r * q[k];
        F[k] += gN;
    }
    free(w);
}

/* solve_benamou_brenier_step (benamou_brenier.py:26-91) */
int oracle_stepA(const double *mu, const double *q, const double *rho0, const double *rhoT,
                 double r, double eps, int Nt, int Nx, int Ny, double rtol, int maxiter,
                 double *phi, int *iters)
{
    size_t N = (size_t)Nt * Nx * Ny;
    double *F = (double *)malloc(N * sizeof(double));
    oracle_rhs(mu, q, rho0, rhoT, r, Nt, Nx, Ny, F);
    int info = oracle_cg(F, r, eps, Nt, Nx, Ny, rtol, maxiter, phi, iters);
    free(F);
    return info;
}

/* stepB (benamou_brenier.py:93-149): projection of (alpha, beta1, beta2) onto
 * K = {alpha + |beta|^2/2 <= 0}; literal transcription of the reference's formulas. */
void oracle_stepB(const double *p, long n, double *q)
{
    for (long i = 0; i < n; i++) {
        double alpha = p[i], beta1 = p[n + i], beta2 = p[2 * n + i];
        if (2 * alpha + beta1 * beta1 + beta2 * beta2 <= 0) {
            q[i] = alpha; q[n + i] = beta1; q[2 * n + i] = beta2;
            continue;
        }
        double rho = sqrt(beta1 * beta1 + beta2 * beta2);
        double theta = atan2(beta2, beta1);
        double alphaH, rhoH, zh;
        if (-32 * pow(alpha + 1, 3) - 108 * rho * rho < 0) {
            double s = 1.0 / 4 * sqrt(2.0) * rho
                     + 1.0 / 6 * sqrt(4.0 / 3 * pow(alpha, 3) + 4 * alpha * alpha + 9.0 / 2 * rho * rho + 4 * alpha + 4.0 / 3);
            double c = pow(s, 1.0 / 3);
            zh = -1.0 / 3 * (alpha + 1) / c;
            zh = zh + c;
            alphaH = -(zh * zh);
            rhoH = sqrt(2.0) * zh;
        } else {
            zh = 2 * sqrt(2.0 / 3) * sqrt(-alpha - 1)
               * cos(1.0 / 3 * acos(pow(3.0 / 2, 3.0 / 2) * rho / pow(-alpha - 1, 3.0 / 2)));
            alphaH = -0.5 * (zh * zh);
            rhoH = zh;
        }
        q[i] = alphaH; q[n + i] = rhoH * cos(theta); q[2 * n + i] = rhoH * sin(theta);
    }
}

/* reconstructTrajectory (utils.py:44-99) for one start pixel; un/vn are [Nt][P]. */
static void trajectory(int xs, int ys, const double *un, const double *vn, int Nx, int Ny, int Nt,
                       double *du, double *dv)
{
    size_t P = (size_t)Nx * Ny;
    double xe = xs, ye = ys;
    for (int n = 0; n < Nt - 1; n++) {
        /* int() truncates toward zero; clamp afterwards.  Clamp in double first so that
         * values outside the int range behave like Python's unbounded int. */
        double tx = trunc(xe), ty = trunc(ye);
        if (tx > Nx - 2) tx = Nx - 2;
        if (tx < 0) tx = 0;
        if (ty > Ny - 2) ty = Ny - 2;
        if (ty < 0) ty = 0;
        int ix = (int)tx, iy = (int)ty;
        double dX = xe - ix, dY = ye - iy;
        double w1 = (1 - dY) * (1 - dX), w2 = dX * (1 - dY), w3 = dY * dX, w4 = (1 - dX) * dY;
        size_t i00 = (size_t)iy * Nx + ix, i01 = i00 + 1, i11 = i00 + Nx + 1, i10 = i00 + Nx;
        const double *u = un + (size_t)n * P, *v = vn + (size_t)n * P;
        xe += (w1 * u[i00] + w2 * u[i01] + w3 * u[i11] + w4 * u[i10]);
        ye += (w1 * v[i00] + w2 * v[i01] + w3 * v[i11] + w4 * v[i10]);
    }
    *du = xe - xs; *dv = ye - ys;
}

/* opticalflow_from_benamoubrenier (utils.py:148-183) with grad(...,'N'), div(...,'D')
 * as passed by solve (benamou_brenier.py:269-271). */
int oracle_flow_from_phi(const double *phi, int Nt, int Nx, int Ny, double *u, double *v, double *m)
{
    if (Nt < 2 || Nx < 2 || Ny < 2) return -1;
    size_t P = (size_t)Nx * Ny;
    double *un = (double *)calloc(2 * (size_t)Nt * P, sizeof(double));
    double *vn = un + (size_t)Nt * P;
    for (int n = 0; n < Nt - 1; n++)
        for (int y = 0; y < Ny; y++)
            for (int x = 0; x < Nx; x++) {
                size_t k = IDX(n, y, x);
                /* grad_1d_central 'N': first and last rows are zero (operators.py:61-63) */
                un[k] = (x == 0 || x == Nx - 1) ? 0.0 : (0.5 * phi[k + 1] - 0.5 * phi[k - 1]);
                vn[k] = (y == 0 || y == Ny - 1) ? 0.0 : (0.5 * phi[k + Nx] - 0.5 * phi[k - Nx]);
            }
    for (int y = 0; y < Ny; y++)
        for (int x = 0; x < Nx; x++)
            trajectory(x, y, un, vn, Nx, Ny, Nt, &u[(size_t)y * Nx + x], &v[(size_t)y * Nx + x]);
    for (int y = 0; y < Ny; y++)
        for (int x = 0; x < Nx; x++) {
            size_t k = (size_t)y * Nx + x;
            /* div with grad_1d_central 'D' (plain central, zero extension), accumulated the way
             * coo_matvec does: x-block entries (x-1, x+1) then y-block entries (y-1, y+1) */
            double s = 0.0;
            if (x > 0) s += -0.5 * u[k - 1];
            if (x + 1 < Nx) s += 0.5 * u[k + 1];
            if (y > 0) s += -0.5 * v[k - Nx];
            if (y + 1 < Ny) s += 0.5 * v[k + Nx];
            m[k] = -s;
        }
    free(un);
    return 0;
}

/* benamou_brenier.solve (benamou_brenier.py:151-271).  crit_trace/cg_iters have room for
 * max_it entries; phi_out (N doubles) may be NULL.  Returns 0, or -1 on bad arguments.
 * cg_rtol/cg_maxiter are 1e-6/1000 in the reference (benamou_brenier.py:85). */
int oracle_solve(const double *rho0, const double *rhoT, int Nt, int Nx, int Ny, double r,
                 double tol, double eps, int max_it, double cg_rtol, int cg_maxiter,
                 double *u, double *v, double *m, double *crit_trace, int *n_outer, int *cg_iters,
                 double *phi_out)
{
    if (Nt < 2 || Nx < 2 || Ny < 2 || max_it < 1) return -1;
    size_t P = (size_t)Nx * Ny, N = P * Nt;
    double *mu = (double *)calloc(3 * N, sizeof(double));
    double *qp = (double *)calloc(3 * N, sizeof(double));
    double *q = (double *)malloc(3 * N * sizeof(double));
    double *gp = (double *)malloc(3 * N * sizeof(double));
    double *pp = (double *)malloc(3 * N * sizeof(double));
    double *phi = (double *)malloc(N * sizeof(double));
    for (int n = 0; n < Nt; n++) {
        double w1 = 1 - (double)n / (Nt - 1), w2 = (double)n / (Nt - 1);
        for (size_t i = 0; i < P; i++) mu[(size_t)n * P + i] = w1 * rho0[i] + w2 * rhoT[i];
    }
    double crit = -1;
    *n_outer = 0;
    for (int it = 0; it < max_it; it++) {
        int iters = 0;
        oracle_stepA(mu, qp, rho0, rhoT, r, eps, Nt, Nx, Ny, cg_rtol, cg_maxiter, phi, &iters);
        cg_iters[it] = iters;
        for (int n = 0; n < Nt; n++)
            for (int y = 0; y < Ny; y++)
                for (int x = 0; x < Nx; x++) {
                    size_t k = IDX(n, y, x);
                    gp[k] = dw(phi, k, P, n, Nt);
                    gp[N + k] = dw(phi, k, 1, x, Nx);
                    gp[2 * N + k] = dw(phi, k, Nx, y, Ny);
                }
        double inv_r = 1.0 / r;
        for (size_t i = 0; i < 3 * N; i++) pp[i] = gp[i] + inv_r * mu[i];
        oracle_stepB(pp, (long)N, q);
        for (size_t i = 0; i < 3 * N; i++) mu[i] = mu[i] + r * (gp[i] - q[i]);
        for (size_t i = 0; i < N; i++) if (!(mu[i] > 0)) mu[i] = mu[i] != mu[i] ? mu[i] : 0.0;
        memcpy(qp, q, 3 * N * sizeof(double));
        double num = 0, den = 0;
        for (size_t i = 0; i < N; i++) {
            double g2 = gp[N + i] * gp[N + i] + gp[2 * N + i] * gp[2 * N + i];
            double res = gp[i] + 0.5 * g2;
            num += mu[i] * fabs(res);
            den += mu[i] * g2;
        }
        double prev = crit;
        crit = sqrt(num / (den + 1e-10));
        crit_trace[it] = crit;
        *n_outer = it + 1;
        if (crit <= tol) break;
        if (prev >= 0 && fabs(prev - crit) < 1e-5) break;
    }
    int rc = oracle_flow_from_phi(phi, Nt, Nx, Ny, u, v, m);
    if (phi_out) memcpy(phi_out, phi, N * sizeof(double));
    free(mu); free(qp); free(q); free(gp); free(pp); free(phi);
    return rc;
}

/* apply_opticalflow (utils.py:186-248).  m may be NULL (no luminosity scaling). */
int oracle_warp_apply(const double *f1, const double *u, const double *v, int w, int h,
                      const double *m, double *out)
{
    size_t P = (size_t)w * h;
    double *g = (double *)malloc(P * sizeof(double));
    for (size_t i = 0; i < P; i++) g[i] = m ? (1 + m[i]) * f1[i] : f1[i];
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            size_t k = (size_t)i * w + j;
            double tI = i - v[k], tJ = j - u[k];
            double dI = tI - trunc(tI), dJ = tJ - trunc(tJ);     /* fractions BEFORE the clamp */
            double w1 = (1 - dI) * (1 - dJ), w2 = dJ * (1 - dI), w3 = dI * dJ, w4 = (1 - dJ) * dI;
            if (tI >= h) tI = h - 1;
            if (tJ >= w) tJ = w - 1;
            if (tI < 0) tI = 0;
            if (tJ < 0) tJ = 0;
            int I = (int)tI, J = (int)tJ;
            int I1 = (int)(tI + 1);                               /* int(tildI+1) */
            double x;
            if (I < h - 1 && J < w - 1) {
                x = w1 * g[(size_t)I * w + J];
                x = x + w2 * g[(size_t)I * w + J + 1];
                x = x + w3 * g[(size_t)I1 * w + J + 1];
                x = x + w4 * g[(size_t)I1 * w + J];
            } else if (I < h - 1 && J == w - 1) {
                x = w1 * g[(size_t)I * w + J];
                x = x + w2 * g[(size_t)I * w + J];
                x = x + w3 * g[(size_t)I1 * w + J];
                x = x + w4 * g[(size_t)I1 * w + J];
            } else if (I == h - 1 && J < w - 1) {
                x = w1 * g[(size_t)I * w + J];
                x = x + w2 * g[(size_t)I * w + J + 1];
                x = x + w3 * g[(size_t)I * w + J + 1];
                x = x + w4 * g[(size_t)I * w + J];
            } else {
                x = w1 * g[(size_t)I * w + J];
                x = x + w2 * g[(size_t)I * w + J];
                x = x + w3 * g[(size_t)I * w + J];
                x = x + w4 * g[(size_t)I * w + J];
            }
            out[k] = x;
        }
    free(g);
    return 0;
}

/* ------------------------------------------------------------------ Gennert-Negahdaripour
 * classical.py:68-111.  g = (fx, fy, -f2); A = diag(alpha,alpha,lambda) (x) (-Lap) + g g^T
 * pointwise; b = -g ft.  Unknown ordering [u | v | m].  */
void oracle_gn_coeffs(const double *f1, const double *f2, int w, int h,
                      double *fx, double *fy, double *ft)
{
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            size_t k = (size_t)i * w + j;
            fx[k] = (j >= 1 && j <= w - 2) ? 0.5 * (f2[k + 1] - f2[k - 1]) : 0.0;
            fy[k] = (i >= 1 && i <= h - 2) ? 0.5 * (f2[k + w] - f2[k - w]) : 0.0;
            ft[k] = f2[k] - f1[k];
        }
}

static inline double neg_lap2(const double *f, size_t k, int j, int i, int w, int h)
{
    return -(lap1(f, k, 1, j, w) + lap1(f, k, (size_t)w, i, h));
}

void oracle_gn_apply(const double *fx, const double *fy, const double *f2, int w, int h,
                     double alpha, double lam, const double *x, double *y)
{
    size_t P = (size_t)w * h;
    const double *u = x, *v = x + P, *m = x + 2 * P;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            size_t k = (size_t)i * w + j;
            y[k] = alpha * neg_lap2(u, k, j, i, w, h) + fx[k] * fx[k] * u[k] + fx[k] * fy[k] * v[k] + (-fx[k] * f2[k]) * m[k];
            y[P + k] = fy[k] * fx[k] * u[k] + alpha * neg_lap2(v, k, j, i, w, h) + fy[k] * fy[k] * v[k] + (-fy[k] * f2[k]) * m[k];
            y[2 * P + k] = (-f2[k] * fx[k]) * u[k] + (-f2[k] * fy[k]) * v[k] + lam * neg_lap2(m, k, j, i, w, h) + f2[k] * f2[k] * m[k];
        }
}

void oracle_gn_rhs(const double *fx, const double *fy, const double *f2, const double *ft,
                   int w, int h, double *b)
{
    size_t P = (size_t)w * h;
    for (size_t k = 0; k < P; k++) { b[k] = -fx[k] * ft[k]; b[P + k] = -fy[k] * ft[k]; b[2 * P + k] = f2[k] * ft[k]; }
}

/* Solve A x = b "exactly" (stand-in for SuperLU): Jacobi-PCG, stop when ||r|| <= rtol ||b||. */
int oracle_gn_solve(const double *f1, const double *f2, int w, int h, double alpha, double lam,
                    double rtol, int maxiter, double *u, double *v, double *m, int *iters)
{
    if (w < 2 || h < 2) return -1;
    size_t P = (size_t)w * h, M = 3 * P;
    double *buf = (double *)malloc((3 * P + 6 * M) * sizeof(double));
    double *fx = buf, *fy = buf + P, *ft = buf + 2 * P;
    double *b = buf + 3 * P, *x = b + M, *r = x + M, *z = r + M, *p = z + M, *q = p + M;
    double *dinv = (double *)malloc(M * sizeof(double));
    oracle_gn_coeffs(f1, f2, w, h, fx, fy, ft);
    oracle_gn_rhs(fx, fy, f2, ft, w, h, b);
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            size_t k = (size_t)i * w + j;
            double deg = (j > 0) + (j < w - 1) + (i > 0) + (i < h - 1);
            dinv[k] = 1.0 / (alpha * deg + fx[k] * fx[k]);
            dinv[P + k] = 1.0 / (alpha * deg + fy[k] * fy[k]);
            dinv[2 * P + k] = 1.0 / (lam * deg + f2[k] * f2[k]);
        }
    memset(x, 0, M * sizeof(double));
    memcpy(r, b, M * sizeof(double));
    double bn = sqrt(dot(b, b, M)), rho_prev = 0;
    int info = maxiter;
    *iters = 0;
    for (int it = 0; it < maxiter; it++) {
        if (sqrt(dot(r, r, M)) <= rtol * bn) { info = 0; break; }
        for (size_t i = 0; i < M; i++) z[i] = dinv[i] * r[i];
        double rho = dot(r, z, M);
        if (it > 0) { double beta = rho / rho_prev; for (size_t i = 0; i < M; i++) p[i] = z[i] + beta * p[i]; }
        else memcpy(p, z, M * sizeof(double));
        oracle_gn_apply(fx, fy, f2, w, h, alpha, lam, p, q);
        double a = rho / dot(p, q, M);
        for (size_t i = 0; i < M; i++) { x[i] += a * p[i]; r[i] -= a * q[i]; }
        rho_prev = rho;
        *iters = it + 1;
    }
    memcpy(u, x, P * sizeof(double)); memcpy(v, x + P, P * sizeof(double)); memcpy(m, x + 2 * P, P * sizeof(double));
    free(buf); free(dinv);
    return info;
}
