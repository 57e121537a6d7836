// cg_kernels.cu -- K2a: the Poisson solve of stepA as ONE persistent cooperative kernel.
//
// Restates scipy.sparse.linalg.cg (x0 = 0, no preconditioner) as called at
// benamou_brenier.py:85 on A = -r L_st + r eps I (benamou_brenier.py:201-203):
//     atol = rtol ||b||;  r = b;  loop k < maxiter:
//         if ||r|| < atol: return (x, 0)
//         rho = r.r;  p = r + (rho/rho_prev) p   (p = r on the first pass)
//         q = A p;  alpha = rho / (p.q);  x += alpha p;  r -= alpha q
//     return (x, maxiter)
// The truncation error of this recurrence at rtol 1e-6 is part of the reference's answer
// (SURVEY.md parity trap #1), so the recurrence is reproduced operation for operation; A p is
// accumulated in csr_matvec's column order (t-1, y-1, x-1, diag, x+1, y+1, t+1).
//
// Streaming variant (this file, cg_stream_kernel): vectors live in global memory (L2-resident up
// to ~10 M cells, HBM beyond).  Two grid-wide barriers per iteration, each fused with the
// all-reduce of the dot product it guards:
//   phase A: p_new = r + beta p_old at the 7 stencil points (p is double-buffered so that
//            neighbours can be recomputed instead of synchronised), q = A p_new, partial p.q
//   phase B: x += alpha p_new, r -= alpha q, partial r.r
// Algorithmic traffic: 11 words per cell per iteration; this
// kernel moves 10 (r, p_old read; p_new, q written; x, p_new, r, q read; x, r written).
#include "foto_kernels.cuh"

namespace foto {

namespace {

struct Coef {            // matrix entries of A, exactly as the reference assembles them
    double off;          // -r * 1.0
    double reps;         // r * eps * 1.0
    double r;
};

__global__ void __launch_bounds__(512, 2) cg_stream_kernel(CgArgs a)
{
    __shared__ double red[128];
    if (a.skip && *a.skip) return;
    const unsigned int P = (unsigned int)a.Nx * (unsigned int)a.Ny;
    const unsigned int N = P * (unsigned int)a.Nt;
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int nth = gridDim.x * blockDim.x;
    const int Nx = a.Nx, Ny = a.Ny, Nt = a.Nt;
    Coef c;
    c.r = a.rcoef; c.off = -a.rcoef * 1.0; c.reps = a.rcoef * a.eps * 1.0;
    unsigned int gen = 0;

    // r = b, x = 0, p_old = 0, ||b||^2
    double acc[1] = {0.0};
    for (unsigned int k = tid; k < N; k += nth) {
        const double bk = a.b[k];
        a.r[k] = bk; a.x[k] = 0.0; a.p0[k] = 0.0;
        acc[0] += bk * bk;
    }
    grid_allreduce<1>(a.sync, gen, acc, red);
    const double bb = acc[0];
    if (*a.sync.error) return;
    if (bb == 0.0) {                                  // scipy: "if bnrm2 == 0: return b, 0"
        if (tid == 0) { a.out[0] = 0; a.out[1] = 0; }
        return;                                       // x already holds zeros (= b)
    }
    const double atol = a.rtol * sqrt(bb);
    double rr = bb, rr_prev = 0.0;
    double *pold = a.p0, *pnew = a.p1;
    int it = 0, info = a.maxiter;
    for (; it < a.maxiter; it++) {
        if (sqrt(rr) < atol) { info = 0; break; }
        const double beta = it > 0 ? rr / rr_prev : 0.0;
        // ---- phase A: one thread per (y, x) column marching through t with p_new(n-1), p_new(n),
        // p_new(n+1) in registers, so p_old and r cross HBM once per iteration even when a plane is
        // far larger than L2 keeps between two visits (the four in-plane neighbours are L1/L2 hits)
        acc[0] = 0.0;
        for (unsigned int i = tid; i < P; i += nth) {
            const int y = (int)(i / (unsigned int)Nx), x = (int)(i - (unsigned int)y * Nx);
            const bool xl = x > 0, xh = x < Nx - 1, yl = y > 0, yh = y < Ny - 1;
            const double dxy = (xl && xh ? -2.0 : -1.0) + (yl && yh ? -2.0 : -1.0);
            double pm = 0.0, pc = pold[i] * beta + a.r[i], pp = pold[P + i] * beta + a.r[P + i];
            for (int n = 0; n < Nt; n++) {
                const unsigned int k = (unsigned int)n * P + i;
                const double Lii = ((n == 0 || n == Nt - 1) ? -1.0 : -2.0) + dxy;
                double s = 0.0;
                if (n > 0) s += c.off * pm;
                if (yl) s += c.off * (pold[k - Nx] * beta + a.r[k - Nx]);
                if (xl) s += c.off * (pold[k - 1] * beta + a.r[k - 1]);
                s += (-c.r * Lii + c.reps) * pc;
                if (xh) s += c.off * (pold[k + 1] * beta + a.r[k + 1]);
                if (yh) s += c.off * (pold[k + Nx] * beta + a.r[k + Nx]);
                if (n < Nt - 1) s += c.off * pp;
                pnew[k] = pc;
                a.q[k] = s;
                acc[0] += pc * s;
                pm = pc; pc = pp;
                if (n + 2 < Nt) pp = pold[k + 2u * P] * beta + a.r[k + 2u * P];
            }
        }
        grid_allreduce<1>(a.sync, gen, acc, red);
        if (*a.sync.error) return;
        const double alpha = rr / acc[0];
        // ---- phase B
        acc[0] = 0.0;
        for (unsigned int k = tid; k < N; k += nth) {
            const double xk = a.x[k] + alpha * pnew[k];
            const double rk = a.r[k] - alpha * a.q[k];
            a.x[k] = xk; a.r[k] = rk;
            acc[0] += rk * rk;
        }
        grid_allreduce<1>(a.sync, gen, acc, red);
        if (*a.sync.error) return;
        rr_prev = rr; rr = acc[0];
        double *t = pold; pold = pnew; pnew = t;
    }
    if (tid == 0) { a.out[0] = it; a.out[1] = info; }
}

}  // namespace

int cg_stream_config(int device, int *grid, int *block)
{
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    int per_sm = 0;
    *block = 512;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_stream_kernel, *block, 0));
    if (per_sm < 1) { set_error("cg_stream_kernel does not fit on an SM"); return FOTO_ERR_CUDA; }
    if (per_sm > 2) per_sm = 2;
    *grid = per_sm * prop.multiProcessorCount;
    if (*grid > kMaxBlocks) *grid = kMaxBlocks;
    return FOTO_OK;
}

int launch_cg_stream(cudaStream_t st, const CgArgs &a, int grid, int block)
{
    CUDA_TRY(cudaMemsetAsync(a.sync.counter, 0, sizeof(unsigned int), st));
    void *args[] = {(void *)&a};
    CUDA_TRY(cudaLaunchCooperativeKernel((void *)cg_stream_kernel, dim3(grid), dim3(block), args, 0, st));
    return FOTO_OK;
}

}  // namespace foto
