// foto_kernels.cu -- FOTO pointwise / stencil kernels (K1, K3, K4, K7, operator apply).
//
// All kernels are fp64, structure-of-arrays, flat index k = n*P + y*Nx + x (x fastest), so
// consecutive threads touch consecutive doubles (fully coalesced 256 B per warp request).
// The file is compiled with -fmad=false: the reference evaluates every product and sum
// separately (numpy / scipy sparse mat-vec), and keeping that rounding makes K1 and K4
// bit-identical to it and keeps the truncated-CG iterates within 1e-11 of the reference.
#include "foto_kernels.cuh"
#include "prox_math.cuh"

namespace foto {

namespace {

constexpr int kThreads = 256;

// "weird" central difference with one-sided Neumann rows (operators.py:33-48), unit spacing
__device__ __forceinline__ double dw(const double *__restrict__ f, unsigned int k, unsigned int st, int i, int n)
{
    if (i == 0) return f[k + st] - f[k];
    if (i == n - 1) return f[k] - f[k - st];
    return 0.5 * f[k + st] - 0.5 * f[k - st];
}

// --------------------------------------------------------------------------- init
__global__ void __launch_bounds__(kThreads) k_init_state(Dims d, const double *__restrict__ rho0,
                                                          const double *__restrict__ rhoT,
                                                          double *__restrict__ mu, double *__restrict__ q)
{
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < d.N; k += stride) {
        unsigned int n = k / d.P, i = k - n * d.P;
        // (1 - n/(Nt-1))*rho0 + (n/(Nt-1))*rhoT     (benamou_brenier.py:193-194)
        double w2 = (double)n / (double)(d.Nt - 1);
        double w1 = 1.0 - w2;
        mu[k] = w1 * rho0[i] + w2 * rhoT[i];
        mu[d.cs + k] = 0.0;
        mu[2u * d.cs + k] = 0.0;
        q[k] = 0.0;
        q[d.cs + k] = 0.0;
        q[2u * d.cs + k] = 0.0;
    }
}

// --------------------------------------------------------------------------- K1
// w = mu - r q evaluated at the stencil points; running sum in scipy coo_matvec order:
// t-block (n-1, n+1), x-block (x-1, x+1), y-block (y-1, y+1); then the rho0/rhoT terms.
__device__ __forceinline__ double wv(const double *__restrict__ mu, const double *__restrict__ q, double r, unsigned int k)
{
    return mu[k] - r * q[k];
}

__device__ __forceinline__ double dw_acc(double s, const double *__restrict__ mu, const double *__restrict__ q,
                                         double r, unsigned int k, unsigned int st, int i, int n)
{
    if (i == 0) { s += -1.0 * wv(mu, q, r, k); s += 1.0 * wv(mu, q, r, k + st); }
    else if (i == n - 1) { s += -1.0 * wv(mu, q, r, k - st); s += 1.0 * wv(mu, q, r, k); }
    else { s += -0.5 * wv(mu, q, r, k - st); s += 0.5 * wv(mu, q, r, k + st); }
    return s;
}

// Measured and rejected (same-box A/B at 1080x1920x16, this kernel: 0.772 of the HBM peak): the 256 threads of a block as an
// 8 x 32 or 4 x 64 tile (rows y-1, y+1 from L1 instead of L2) 0.72-0.74; stencil points and weights as per-column constants with
// two planes per trip 0.753, one 0.63, four 0.62; a thread owning a strip of four rows with the y taps from one register
// window of six rows (1.5 instead of 2 requests per cell for that component, bit-identical F) 0.51 (profiles/r2_k1_rows.log).
// A grid-stride kernel that only reads six streams and writes one reaches 1.03 of the copy peak (tools/ubench_streams.cu);
// this kernel pulls 70 B per cell through L2 for its halo taps on top of the 56 B of HBM traffic, but neither variant that
// cuts those requests (tiles, row strips) came out ahead.  The kernel lives on threads in flight: with 6 / 5 / 4 instead of 8
// resident blocks per SM (40 / 48 / 52 registers, no spills instead of 60 bytes) it drops to 0.737 / 0.690 / 0.611
// (profiles/r2_k1_rows.log), yet it is not waiting for DRAM latency either: prefetch.global.L2 of the taps 1 / 2 / 3 planes
// ahead (one request per 128-byte line) costs 0.690 / 0.668 / 0.626 -- extra requests hurt, so the limit is the request rate
// through L1 / L2 (10 loads per cell for 6 new words).  Taking the x taps from the neighbouring lanes by warp shuffle (8.1 loads
// per cell, bit-identical F) costs more than it saves: 0.70.  What is left is staging the tiles once (TMA, as K3 does).
// One thread per (y, x) column marching through t: w_t = (mu - r q)_rho of the planes n-1, n, n+1
// stays in registers, so every word of mu and q is read once from HBM even when a plane (16 MB at
// 1080x1920) is far larger than what L2 keeps between two visits.
__global__ void __launch_bounds__(kThreads, 8) k_rhs(Dims d, const double *__restrict__ mu, const double *__restrict__ q,
                                                   const double *__restrict__ rho0, const double *__restrict__ rhoT,
                                                   double r, double *__restrict__ F)
{
    if (d.skip && *d.skip) return;
    const unsigned int stride = gridDim.x * blockDim.x;
    const double *mu1 = mu + d.cs, *q1 = q + d.cs, *mu2 = mu + 2u * d.cs, *q2 = q + 2u * d.cs;
    auto has = [&](int nl) { const int gn = d.n0 + nl; return gn >= 0 && gn < d.gNt; };   // plane exists globally
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.P; i += stride) {
        const int y = (int)(i / (unsigned int)d.Nx), x = (int)(i - (unsigned int)y * d.Nx);
        double w_m = has(-1) ? mu[(long long)i - (long long)d.P] - r * q[(long long)i - (long long)d.P] : 0.0;
        double w_c = wv(mu, q, r, i), w_p = has(1) ? wv(mu, q, r, d.P + i) : 0.0;
        for (int n = 0; n < d.Nt; n++) {
            const unsigned int k = (unsigned int)n * d.P + i;
            const int gn = d.n0 + n;
            double s = 0.0;
            if (gn == 0) { s += -1.0 * w_c; s += 1.0 * w_p; }
            else if (gn == d.gNt - 1) { s += -1.0 * w_m; s += 1.0 * w_c; }
            else { s += -0.5 * w_m; s += 0.5 * w_p; }
            s = dw_acc(s, mu1, q1, r, k, 1u, x, d.Nx);
            s = dw_acc(s, mu2, q2, r, k, (unsigned int)d.Nx, y, d.Ny);
            if (gn == 0) s -= (rho0[i] - mu[k] + r * q[k]);
            if (gn == d.gNt - 1) s += (rhoT[i] - mu[k] + r * q[k]);
            F[k] = s;
            w_m = w_c; w_c = w_p;
            if (n + 2 <= d.Nt && has(n + 2)) w_p = wv(mu, q, r, k + 2u * d.P);   // local planes -1 .. Nt only
        }
    }
}

// --------------------------------------------------------------------------- stepB (project_K: prox_math.cuh)
__global__ void __launch_bounds__(kThreads) k_stepB(unsigned int N, const double *__restrict__ p, double *__restrict__ q)
{
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
        double qa, qb1, qb2;
        project_K(p[k], p[N + k], p[2u * N + k], qa, qb1, qb2);
        q[k] = qa; q[N + k] = qb1; q[2u * N + k] = qb2;
    }
}

// --------------------------------------------------------------------------- K3
// gradPhi = grad_st phi (registers only); p = gradPhi + mu/r; q = stepB(p);
// mu += r (gradPhi - q); mu_rho = max(mu_rho, 0); criterion partial sums with the updated mu.
// Threads march through t with phi(n-1), phi(n), phi(n+1) in registers: phi is read once from HBM
// (a flat sweep re-reads it three times at HD size).
constexpr int kTChunk = 1 << 20;    // one chunk: splitting t across threads (4 planes each) measured slower at 1080x1920x16
__global__ void __launch_bounds__(kThreads, 4) k_prox_dual(Dims d, const double *__restrict__ phi, double *__restrict__ mu,
                                                         double *__restrict__ q, double r, double inv_r,
                                                         double *__restrict__ partials)
{
    __shared__ double red[64];
    if (d.skip && *d.skip) return;
    double acc[2] = {0.0, 0.0};
    const unsigned int stride = gridDim.x * blockDim.x;
    // work item = (chunk of kTChunk time planes, column): marching keeps phi(n-1), phi(n), phi(n+1) in
    // registers (one extra phi word per chunk end), chunking keeps enough threads in flight when Nt is large
    const unsigned int nchunk = ((unsigned int)d.Nt + kTChunk - 1) / kTChunk, items = nchunk * d.P;
    for (unsigned int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += stride) {
        const unsigned int chunk = it / d.P, i = it - chunk * d.P;
        const int y = (int)(i / (unsigned int)d.Nx), x = (int)(i - (unsigned int)y * d.Nx);
        const int nb = (int)(chunk * kTChunk), ne = min(nb + kTChunk, d.Nt);
        const unsigned int kb = (unsigned int)nb * d.P + i;
        auto has = [&](int nl) { const int gn = d.n0 + nl; return gn >= 0 && gn < d.gNt; };   // plane exists globally
        double p_m = has(nb - 1) ? phi[(long long)kb - (long long)d.P] : 0.0, p_c = phi[kb], p_p = has(nb + 1) ? phi[kb + d.P] : 0.0;
        // software pipeline: the three mu words of plane n+1 are requested before the projection of
        // plane n is computed (the kernel is latency-bound: 4 HBM words per cell and iteration in flight)
        double n0 = mu[kb], n1 = mu[d.cs + kb], n2 = mu[2u * d.cs + kb];
        for (int n = nb; n < ne; n++) {
            const unsigned int k = (unsigned int)n * d.P + i;
            const int gn = d.n0 + n;
            const double gt = gn == 0 ? p_p - p_c : (gn == d.gNt - 1 ? p_c - p_m : 0.5 * p_p - 0.5 * p_m);
            const double gx = dw(phi, k, 1u, x, d.Nx);
            const double gy = dw(phi, k, (unsigned int)d.Nx, y, d.Ny);
            const double m0 = n0, m1 = n1, m2 = n2;
            p_m = p_c; p_c = p_p;
            if (n + 2 <= d.Nt && has(n + 2)) p_p = phi[k + 2u * d.P];      // local planes -1 .. Nt only (Nt = halo of a slab)
            if (n + 1 < ne) { n0 = mu[k + d.P]; n1 = mu[d.cs + k + d.P]; n2 = mu[2u * d.cs + k + d.P]; }
            double qa, qb1, qb2;
            project_K(gt + inv_r * m0, gx + inv_r * m1, gy + inv_r * m2, qa, qb1, qb2);
            q[k] = qa; q[d.cs + k] = qb1; q[2u * d.cs + k] = qb2;
            double rho = m0 + r * (gt - qa);
            rho = fmax(rho, 0.0);
            mu[k] = rho;
            mu[d.cs + k] = m1 + r * (gx - qb1);
            mu[2u * d.cs + k] = m2 + r * (gy - qb2);
            const double g2 = gx * gx + gy * gy;
            const double res = gt + 0.5 * g2;
            acc[0] += rho * fabs(res);
            acc[1] += rho * g2;
        }
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = acc[0]; partials[2 * blockIdx.x + 1] = acc[1]; }
}

__global__ void __launch_bounds__(kThreads) k_crit_final(const double *__restrict__ partials, int blocks, double *__restrict__ out2)
{
    __shared__ double red[64];
    double acc[2] = {0.0, 0.0};
    for (int b = threadIdx.x; b < blocks; b += blockDim.x) { acc[0] += partials[2 * b]; acc[1] += partials[2 * b + 1]; }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) { out2[0] = acc[0]; out2[1] = acc[1]; }
}

// Outer-loop bookkeeping on the device: crit = sqrt(sum rho |res| / (sum rho |grad_x phi|^2 + 1e-10)) and the
// stopping rule "crit <= tol or |crit_prev - crit| < 1e-5 (from the second iteration on)" (benamou_brenier.py:251-258),
// so that the host can enqueue the next iteration without waiting for this one (sqrt and the division are IEEE
// correctly rounded on both sides: same bits as the host evaluation).
__global__ void __launch_bounds__(kThreads) k_outer_decide(const double *__restrict__ partials, int blocks, double *__restrict__ sums2,
                                                            const int *__restrict__ cg_out, OuterState *state, OuterTrace tr,
                                                            int it, double tol, int max_it)
{
    __shared__ double red[64];
    if (state->done) return;                             // uniform: written below only after the block-wide sums
    double acc[2] = {0.0, 0.0};
    for (int b = threadIdx.x; b < blocks; b += blockDim.x) { acc[0] += partials[2 * b]; acc[1] += partials[2 * b + 1]; }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
        sums2[0] = acc[0]; sums2[1] = acc[1];
        const double crit = sqrt(acc[0] / (acc[1] + 1e-10));
        tr.crit[it] = crit; tr.cg_iters[it] = cg_out[0]; tr.cg_info[it] = cg_out[1];
        const bool stop = crit <= tol || (it > 0 && fabs(state->crit_prev - crit) < 1e-5) || it + 1 >= max_it;
        state->crit_prev = crit;
        state->n_outer = it + 1;
        if (stop) state->done = 1;
    }
}

// --------------------------------------------------------------------------- K4
// Velocity taps un = Dc_x phi_n, vn = Dc_y phi_n (grad_1d_central 'N': zero on the first and
// last column/row, operators.py:61-63) are recomputed from phi at each bilinear tap; phi of one
// time slice is L2-resident, so the 16 gathers per step cost no HBM traffic.
__device__ __forceinline__ double tap_u(const double *__restrict__ ph, unsigned int i, int x, int Nx)
{
    return (x == 0 || x == Nx - 1) ? 0.0 : (0.5 * ph[i + 1] - 0.5 * ph[i - 1]);
}
__device__ __forceinline__ double tap_v(const double *__restrict__ ph, unsigned int i, int y, int Ny, int Nx)
{
    return (y == 0 || y == Ny - 1) ? 0.0 : (0.5 * ph[i + Nx] - 0.5 * ph[i - Nx]);
}

__global__ void __launch_bounds__(kThreads) k_trajectories(Dims d, const double *__restrict__ phi,
                                                            double *__restrict__ u, double *__restrict__ v)
{
    const unsigned int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= d.P) return;
    const int ys = (int)(i0 / (unsigned int)d.Nx), xs = (int)(i0 - (unsigned int)ys * d.Nx);
    double xe = (double)xs, ye = (double)ys;
    for (int n = 0; n < d.Nt - 1; n++) {
        // int() truncation toward zero, then clamp to [0, N-2] (utils.py:63-68); the clamp is
        // done in double so that huge / negative excursions behave like Python's unbounded int
        double tx = trunc(xe), ty = trunc(ye);
        tx = fmin(tx, (double)(d.Nx - 2)); tx = fmax(tx, 0.0);
        ty = fmin(ty, (double)(d.Ny - 2)); ty = fmax(ty, 0.0);
        const int ix = (int)tx, iy = (int)ty;
        const double dX = xe - (double)ix, dY = ye - (double)iy;
        const double w1 = (1.0 - dY) * (1.0 - dX), w2 = dX * (1.0 - dY), w3 = dY * dX, w4 = (1.0 - dX) * dY;
        const double *ph = phi + (size_t)n * d.P;
        const unsigned int i00 = (unsigned int)iy * d.Nx + ix;
        const double u00 = tap_u(ph, i00, ix, d.Nx), u01 = tap_u(ph, i00 + 1, ix + 1, d.Nx);
        const double u11 = tap_u(ph, i00 + d.Nx + 1, ix + 1, d.Nx), u10 = tap_u(ph, i00 + d.Nx, ix, d.Nx);
        const double v00 = tap_v(ph, i00, iy, d.Ny, d.Nx), v01 = tap_v(ph, i00 + 1, iy, d.Ny, d.Nx);
        const double v11 = tap_v(ph, i00 + d.Nx + 1, iy + 1, d.Ny, d.Nx), v10 = tap_v(ph, i00 + d.Nx, iy + 1, d.Ny, d.Nx);
        xe += (w1 * u00 + w2 * u01 + w3 * u11 + w4 * u10);
        ye += (w1 * v00 + w2 * v01 + w3 * v11 + w4 * v10);
    }
    u[i0] = xe - (double)xs;
    v[i0] = ye - (double)ys;
}

// m = -div(u, v) with grad_1d_central 'D' (zero extension), summed in coo_matvec order
__global__ void __launch_bounds__(kThreads) k_luminosity(Dims d, const double *__restrict__ u, const double *__restrict__ v,
                                                          double *__restrict__ m)
{
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.P) return;
    const int y = (int)(k / (unsigned int)d.Nx), x = (int)(k - (unsigned int)y * d.Nx);
    double s = 0.0;
    if (x > 0) s += -0.5 * u[k - 1];
    if (x + 1 < d.Nx) s += 0.5 * u[k + 1];
    if (y > 0) s += -0.5 * v[k - d.Nx];
    if (y + 1 < d.Ny) s += 0.5 * v[k + d.Nx];
    m[k] = -s;
}

// --------------------------------------------------------------------------- K7
__global__ void __launch_bounds__(kThreads) k_scale_lum(unsigned int P, const double *__restrict__ f1,
                                                         const double *__restrict__ m, double *__restrict__ g)
{
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < P) g[k] = m ? (1.0 + m[k]) * f1[k] : f1[k];
}

// Backward bilinear warp with the reference's quirks (utils.py:205-247): fractions are taken
// BEFORE clamping, truncation is toward zero, the lower row index is int(tildI + 1), and the
// four taps are accumulated left to right.
__global__ void __launch_bounds__(kThreads) k_warp(int w, int h, const double *__restrict__ g, const double *__restrict__ u,
                                                    const double *__restrict__ v, double *__restrict__ out)
{
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (unsigned int)w * (unsigned int)h) return;
    const int i = (int)(k / (unsigned int)w), j = (int)(k - (unsigned int)i * w);
    double tI = (double)i - v[k], tJ = (double)j - u[k];
    const double dI = tI - trunc(tI), dJ = tJ - trunc(tJ);
    const double w1 = (1.0 - dI) * (1.0 - dJ), w2 = dJ * (1.0 - dI), w3 = dI * dJ, w4 = (1.0 - dJ) * dI;
    if (tI >= (double)h) tI = (double)(h - 1);
    if (tJ >= (double)w) tJ = (double)(w - 1);
    if (tI < 0.0) tI = 0.0;
    if (tJ < 0.0) tJ = 0.0;
    const int I = (int)tI, J = (int)tJ, I1 = (int)(tI + 1.0);
    const bool right = (J == w - 1), bottom = (I == h - 1);
    const size_t r0 = (size_t)I * w, r1 = (size_t)(bottom ? I : I1) * w;
    const int J1 = right ? J : J + 1;
    double x = w1 * g[r0 + J];
    x = x + w2 * g[r0 + J1];
    x = x + w3 * g[r1 + J1];
    x = x + w4 * g[r1 + J];
    out[k] = x;
}

// --------------------------------------------------------------------------- operators
__global__ void __launch_bounds__(kThreads) k_axis_apply(const double *__restrict__ in, double *__restrict__ out,
                                                          const double *__restrict__ lo, const double *__restrict__ di,
                                                          const double *__restrict__ up, int transpose,
                                                          unsigned int stride, int len, unsigned int total, int accumulate)
{
    const unsigned int gs = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gs) {
        const int c = (int)((k / stride) % (unsigned int)len);
        double s = 0.0;
        if (!transpose) {
            if (c > 0) s += lo[c] * in[k - stride];
            s += di[c] * in[k];
            if (c + 1 < len) s += up[c] * in[k + stride];
        } else {
            if (c > 0) s += up[c - 1] * in[k - stride];
            s += di[c] * in[k];
            if (c + 1 < len) s += lo[c + 1] * in[k + stride];
        }
        out[k] = accumulate ? out[k] + s : s;
    }
}

inline int blocks_for(unsigned int n, int cap = 148 * 16)
{
    unsigned int b = (n + kThreads - 1) / kThreads;
    return (int)(b < (unsigned int)cap ? (b ? b : 1) : (unsigned int)cap);
}

}  // namespace

void launch_init_state(cudaStream_t st, Dims d, const double *rho0, const double *rhoT, double *mu, double *q)
{
    k_init_state<<<blocks_for(d.N), kThreads, 0, st>>>(d, rho0, rhoT, mu, q);
}

void launch_rhs(cudaStream_t st, Dims d, const double *mu, const double *q, const double *rho0, const double *rhoT,
                double r, double *F)
{
    k_rhs<<<blocks_for(d.P), kThreads, 0, st>>>(d, mu, q, rho0, rhoT, r, F);
}

int launch_prox_dual(cudaStream_t st, Dims d, const double *phi, double *mu, double *q, double r, double *partials,
                     int max_blocks, int num_sms, int *variant)
{
    if (prox_tma_eligible(d, phi, mu, q)) {              // TMA-staged variant (prox_tma.cu): whole volumes, even Nx
        int blocks = 0;
        if (variant) *variant = 1;
        if (launch_prox_dual_tma(st, d, phi, mu, q, r, partials, max_blocks, num_sms, &blocks) == FOTO_OK) return blocks;
        return -1;
    }
    if (variant) *variant = 0;
    int blocks = blocks_for(d.P, max_blocks);
    k_prox_dual<<<blocks, kThreads, 0, st>>>(d, phi, mu, q, r, 1.0 / r, partials);
    return blocks;
}

void launch_crit_final(cudaStream_t st, const double *partials, int blocks, double *out2)
{
    k_crit_final<<<1, kThreads, 0, st>>>(partials, blocks, out2);
}

void launch_outer_decide(cudaStream_t st, const double *partials, int blocks, double *crit_sums2, const int *cg_out,
                         OuterState *state, OuterTrace trace, int it, double tol, int max_it)
{
    k_outer_decide<<<1, kThreads, 0, st>>>(partials, blocks, crit_sums2, cg_out, state, trace, it, tol, max_it);
}

void launch_stepB(cudaStream_t st, unsigned int N, const double *p, double *q)
{
    k_stepB<<<blocks_for(N), kThreads, 0, st>>>(N, p, q);
}

void launch_flow(cudaStream_t st, Dims d, const double *phi, double *u, double *v, double *m)
{
    const int blocks = (int)((d.P + kThreads - 1) / kThreads);
    k_trajectories<<<blocks, kThreads, 0, st>>>(d, phi, u, v);
    k_luminosity<<<blocks, kThreads, 0, st>>>(d, u, v, m);
}

void launch_warp(cudaStream_t st, int w, int h, const double *f1, const double *u, const double *v,
                 const double *m_or_null, double *g, double *out)
{
    const unsigned int P = (unsigned int)w * (unsigned int)h;
    const int blocks = (int)((P + kThreads - 1) / kThreads);
    k_scale_lum<<<blocks, kThreads, 0, st>>>(P, f1, m_or_null, g);
    k_warp<<<blocks, kThreads, 0, st>>>(w, h, g, u, v, out);
}

void launch_axis_apply(cudaStream_t st, const double *in, double *out, const double *lo, const double *di,
                       const double *up, int transpose, unsigned int stride, int len, unsigned int total, int accumulate)
{
    k_axis_apply<<<blocks_for(total), kThreads, 0, st>>>(in, out, lo, di, up, transpose, stride, len, total, accumulate);
}

}  // namespace foto

// ------------------------------------------------------------------------------- next-tier rows
// (SURVEY.md section 8f): .flo egress packing and endpoint/angular error metrics on the device.
namespace foto {
namespace {

// Middlebury .flo payload: float32 (u, v) interleaved, row-major (utils.py:285-292)
__global__ void __launch_bounds__(256) k_pack_flo(unsigned int n, const double *__restrict__ u,
                                                   const double *__restrict__ v, float2 *__restrict__ out)
{
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        out[k] = make_float2((float)u[k], (float)v[k]);
}

// EE = sqrt((u-uGT)^2 + (v-vGT)^2), kept when EE <= 50 (utils.py:308-312);
// AE = acos((1 + u uGT + v vGT) / (sqrt(1+u^2+v^2) sqrt(1+uGT^2+vGT^2))), kept when not NaN (utils.py:331-335).
// out6 = [sum EE, sum EE^2, count EE, sum AE, sum AE^2, count AE]
__global__ void __launch_bounds__(256) k_flow_metrics(unsigned int n, const double *__restrict__ u, const double *__restrict__ v,
                                                       const double *__restrict__ ug, const double *__restrict__ vg,
                                                       double *__restrict__ partials)
{
    __shared__ double red[32 * 4];
    double acc[4] = {0.0, 0.0, 0.0, 0.0}, cnt[2] = {0.0, 0.0};
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const double du = u[k] - ug[k], dv = v[k] - vg[k];
        const double ee = sqrt(du * du + dv * dv);
        if (ee <= 50.0) { acc[0] += ee; acc[1] += ee * ee; cnt[0] += 1.0; }
        const double ae = acos((1.0 + u[k] * ug[k] + v[k] * vg[k]) /
                               (sqrt(1.0 + u[k] * u[k] + v[k] * v[k]) * sqrt(1.0 + ug[k] * ug[k] + vg[k] * vg[k])));
        if (ae == ae) { acc[2] += ae; acc[3] += ae * ae; cnt[1] += 1.0; }
    }
    block_sum<4>(acc, red);
    block_sum<2>(cnt, red);
    if (threadIdx.x == 0) {
        double *p = partials + 6 * blockIdx.x;
        p[0] = acc[0]; p[1] = acc[1]; p[2] = cnt[0]; p[3] = acc[2]; p[4] = acc[3]; p[5] = cnt[1];
    }
}

__global__ void __launch_bounds__(256) k_sum6(const double *__restrict__ partials, int blocks, double *__restrict__ out6)
{
    __shared__ double red[32 * 4];
    double a[4] = {0, 0, 0, 0}, b[2] = {0, 0};
    for (int i = threadIdx.x; i < blocks; i += blockDim.x) {
        const double *p = partials + 6 * i;
        a[0] += p[0]; a[1] += p[1]; b[0] += p[2]; a[2] += p[3]; a[3] += p[4]; b[1] += p[5];
    }
    block_sum<4>(a, red);
    block_sum<2>(b, red);
    if (threadIdx.x == 0) { out6[0] = a[0]; out6[1] = a[1]; out6[2] = b[0]; out6[3] = a[2]; out6[4] = a[3]; out6[5] = b[1]; }
}

// utils.openGrayscaleImage: f.flatten() / 255 (utils.py:42) -- an IEEE division of two exact integers, as numpy's
__global__ void __launch_bounds__(256) k_ingest_u8(unsigned int n, const unsigned char *__restrict__ in, double *__restrict__ out)
{
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = (double)in[k] / 255.0;
}

// utils.IE (utils.py:354): sum (255 I - 255 IGT)^2
__global__ void __launch_bounds__(256) k_ie_partial(unsigned int n, const double *__restrict__ a, const double *__restrict__ b,
                                                     double *__restrict__ partials)
{
    __shared__ double red[32];
    double acc[1] = {0.0};
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const double d = 255.0 * a[k] - 255.0 * b[k];
        acc[0] += d * d;
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
}
__global__ void __launch_bounds__(256) k_sum1(const double *__restrict__ partials, int blocks, double *__restrict__ out1)
{
    __shared__ double red[32];
    double acc[1] = {0.0};
    for (int i = threadIdx.x; i < blocks; i += blockDim.x) acc[0] += partials[i];
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) out1[0] = acc[0];
}

// Time-slab transpose, one pass: element (l, y, x) of this rank's L planes sits at
//   off_g * L * Nx + (l * rows_g + (y - y0_g)) * Nx + x,   g = owner of row y, y0_g = g Ny / world (slab.split),
// in the all-to-all buffer, i.e. the block sent to / received from rank g is contiguous.  pack: planes -> buffer;
// unpack: buffer -> planes.
__global__ void __launch_bounds__(256) k_slab_pack(int L, int Ny, int Nx, int world, const double *__restrict__ pin,
                                                    double *__restrict__ bout, double *__restrict__ pout, const double *__restrict__ bin)
{
    const size_t total = (size_t)L * Ny * Nx, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        const int x = (int)(k % Nx);
        const size_t ly = k / Nx;
        const int y = (int)(ly % Ny), l = (int)(ly / Ny);
        int g = (int)(((long long)(y + 1) * world - 1) / Ny);          // owner of row y: g Ny / world <= y < (g+1) Ny / world
        while ((long long)g * Ny / world > y) g--;
        while ((long long)(g + 1) * Ny / world <= y) g++;
        const int y0 = (int)((long long)g * Ny / world), rows = (int)((long long)(g + 1) * Ny / world) - y0;
        const size_t b = ((size_t)y0 * L + (size_t)l * rows + (size_t)(y - y0)) * Nx + x;
        if (pin) bout[b] = pin[k]; else pout[k] = bin[b];
    }
}

}  // namespace

void launch_slab_pack(cudaStream_t st, int L, int Ny, int Nx, int world, const double *planes_in, double *buf_out,
                      double *planes_out, const double *buf_in)
{
    const size_t total = (size_t)L * Ny * Nx;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_slab_pack<<<blocks, 256, 0, st>>>(L, Ny, Nx, world, planes_in, buf_out, planes_out, buf_in);
}

void launch_ingest_u8(cudaStream_t st, unsigned int n, const unsigned char *in, double *out)
{
    k_ingest_u8<<<blocks_for(n), 256, 0, st>>>(n, in, out);
}

void launch_ie_sumsq(cudaStream_t st, unsigned int n, const double *a, const double *b, double *partials, double *out1)
{
    const int blocks = blocks_for(n, 148 * 8);
    k_ie_partial<<<blocks, 256, 0, st>>>(n, a, b, partials);
    k_sum1<<<1, 256, 0, st>>>(partials, blocks, out1);
}

void launch_pack_flo(cudaStream_t st, unsigned int n, const double *u, const double *v, float *out)
{
    k_pack_flo<<<blocks_for(n), 256, 0, st>>>(n, u, v, (float2 *)out);
}

// partials: 6 * 1184 doubles
void launch_flow_metrics(cudaStream_t st, unsigned int n, const double *u, const double *v, const double *ug,
                         const double *vg, double *partials, double *out6)
{
    const int blocks = blocks_for(n, 148 * 8);
    k_flow_metrics<<<blocks, 256, 0, st>>>(n, u, v, ug, vg, partials);
    k_sum6<<<1, 256, 0, st>>>(partials, blocks, out6);
}

}  // namespace foto
