// cg_slab.cu -- K2a for ONE volume split into time slabs over several GPUs: the reference's truncated CG
// (benamou_brenier.py:85, scipy cg with rtol 1e-6, maxiter 1000, x0 = 0) with the recurrences of cg_kernels.cu, cut into
// the pieces between which the ranks have to talk:
//
//     init      r = b, x = 0, p = 0, partial b.b                                   -> all-reduce
//     phase A   stop test and beta from the all-reduced r.r (every thread, same bits); p_new = p_old beta + r on the owned
//               planes AND on the two halo planes (same arithmetic as the owner, so p never travels), q = A p_new in
//               csr_matvec order, partial p.q                                        -> all-reduce
//     phase B   alpha = r.r / p.q; x += alpha p_new, r -= alpha q, partial r.r       -> all-reduce; halo planes of r -> neighbours
// The scalars, the iteration count, scipy's info and a `done` flag live in a small device-side state block that the last
// block of a kernel to finish updates (after every block has read it): the host never waits for a scalar.
//
// The collectives (a one-word all-reduce after each phase, one boundary plane of r per neighbour and iteration) are
// issued by the caller (foto_b200/slab.py, torch.distributed over NCCL) on the same stream.  Every kernel of an iteration
// returns at once when the device-side `done` flag is up, so the host enqueues iterations ahead and looks at the flag
// every few iterations only.  Dot products are summed per block in a fixed order and by the last block to finish, i.e.
// reproducibly; their order differs from the one-GPU kernels', so the iteration count can differ by one in the rare
// solve whose last residual lands within 1e-4 of the threshold.
#include "foto_kernels.cuh"

namespace foto {

namespace {

constexpr int kThreads = 256, kSlabBlocks = 148 * 8;

// state words (doubles): 0 value in flight (local partial in, all-reduced total out), 1 r.r of the current pass, 3 atol, 4 done,
// 5 iterations, 6 info, 10 ticket (as an integer word)
enum { S_VAL = 0, S_RR, S_RRPREV, S_ATOL, S_DONE, S_ITERS, S_INFO, S_BB, S_BETA, S_ALPHA, S_TICKET, S_WORDS = 16 };   // RRPREV, BB, BETA, ALPHA: unused

// block partials -> state[S_VAL], summed in block order by the last block to finish, which also applies `upd` to the state
// (every block has read the state by then)
struct Upd { int n; int idx[4]; double val[4]; };

__device__ void publish_sum(double acc, double *partials, double *state, double *red, const Upd &upd, bool have_sum)
{
    double v[1] = {acc};
    if (have_sum) block_sum<1>(v, red);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        if (have_sum) partials[blockIdx.x] = v[0];
        __threadfence();
        is_last = atomicAdd((unsigned int *)(state + S_TICKET), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s[1] = {0.0};
        if (have_sum) {
            for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) s[0] += __ldcg(partials + b);
            block_sum<1>(s, red);
        }
        if (threadIdx.x == 0) {
            for (int i = 0; i < upd.n; i++) state[upd.idx[i]] = upd.val[i];
            if (have_sum) state[S_VAL] = s[0];
            *(unsigned int *)(state + S_TICKET) = 0u;
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_init(unsigned int N, unsigned int P, const double *__restrict__ b, double *__restrict__ x,
                                                    double *__restrict__ r, double *__restrict__ p, double *partials, double *state)
{
    __shared__ double red[64];
    double acc = 0.0;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < N + 2u * P; k += stride) {      // r, p: planes -1 .. nloc
        double bk = 0.0;
        if (k >= P && k < N + P) { bk = b[k - P]; x[k - P] = 0.0; acc += bk * bk; }
        r[k] = bk; p[k] = 0.0;
    }
    publish_sum(acc, partials, state, red, Upd{0, {}, {}}, true);
}

// after maxiter iterations without convergence: scipy returns (x, maxiter)
__global__ void k_finish(double *state, int maxiter)
{
    if (state[S_DONE] != 0.0) return;
    state[S_DONE] = 1.0; state[S_ITERS] = (double)maxiter; state[S_INFO] = (double)maxiter;
}

// phase A.  r, pold, pnew: [nloc + 2][P] with planes -1 and nloc as halos (pointers to plane 0); q: [nloc][P]
__global__ void __launch_bounds__(kThreads) k_phase_a(int gNt, int n0, int nloc, int Ny, int Nx, double rcoef, double eps, double rtol, int it,
                                                       const double *__restrict__ r, const double *__restrict__ pold,
                                                       double *__restrict__ pnew, double *__restrict__ q, double *partials, double *state)
{
    __shared__ double red[64];
    if (state[S_DONE] != 0.0) return;
    // scipy's pass `it`: rr = r.r (b.b on the first pass) is the value the ranks have just all-reduced
    const double rr = state[S_VAL], rr_prev = state[S_RR];
    const double atol = it == 0 ? rtol * sqrt(rr) : state[S_ATOL];
    if ((it == 0 && rr == 0.0) || sqrt(rr) < atol) {     // "if bnrm2 == 0: return b, 0" / "if ||r|| < atol: return x, 0"
        publish_sum(0.0, partials, state, red, Upd{3, {S_DONE, S_ITERS, S_INFO}, {1.0, (double)it, 0.0}}, false);
        return;
    }
    const double beta = it > 0 ? rr / rr_prev : 0.0;
    const double off = -rcoef * 1.0, reps = rcoef * eps * 1.0;
    const long long P = (long long)Nx * Ny;
    const bool lo = n0 > 0, hi = n0 + nloc < gNt;        // a neighbour slab exists below / above
    double acc = 0.0;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned int)P; i += stride) {
        const int y = (int)(i / (unsigned int)Nx), x = (int)(i - (unsigned int)y * Nx);
        const bool xl = x > 0, xh = x < Nx - 1, yl = y > 0, yh = y < Ny - 1;
        const double dxy = (xl && xh ? -2.0 : -1.0) + (yl && yh ? -2.0 : -1.0);
        double pm = 0.0;
        if (lo) { pm = pold[(long long)i - P] * beta + r[(long long)i - P]; pnew[(long long)i - P] = pm; }
        double pc = pold[i] * beta + r[i];
        double pp = (nloc > 1 || hi) ? pold[P + i] * beta + r[P + i] : 0.0;
        for (int n = 0; n < nloc; n++) {
            const long long k = (long long)n * P + i;
            const int gn = n0 + n;
            const double Lii = ((gn == 0 || gn == gNt - 1) ? -1.0 : -2.0) + dxy;
            double s = 0.0;
            if (gn > 0) s += off * pm;
            if (yl) s += off * (pold[k - Nx] * beta + r[k - Nx]);
            if (xl) s += off * (pold[k - 1] * beta + r[k - 1]);
            s += (-rcoef * Lii + reps) * pc;
            if (xh) s += off * (pold[k + 1] * beta + r[k + 1]);
            if (yh) s += off * (pold[k + Nx] * beta + r[k + Nx]);
            if (gn < gNt - 1) s += off * pp;
            pnew[k] = pc;
            q[k] = s;
            acc += pc * s;
            pm = pc; pc = pp;
            if (n + 2 < nloc || (n + 2 == nloc && hi)) pp = pold[k + 2 * P] * beta + r[k + 2 * P];
        }
        if (hi) pnew[(long long)nloc * P + i] = pc;       // after the loop pc holds the plane nloc
    }
    publish_sum(acc, partials, state, red, Upd{2, {S_RR, S_ATOL}, {rr, atol}}, true);       // rr: alpha of phase B, rr_prev of the next pass
}

__global__ void __launch_bounds__(kThreads) k_phase_b(unsigned int N, const double *__restrict__ p, const double *__restrict__ q,
                                                       double *__restrict__ x, double *__restrict__ r, double *partials, double *state)
{
    __shared__ double red[64];
    if (state[S_DONE] != 0.0) return;
    const double alpha = state[S_RR] / state[S_VAL];     // r.r / p.q (just all-reduced)
    double acc = 0.0;
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
        const double xk = x[k] + alpha * p[k];
        const double rk = r[k] - alpha * q[k];
        x[k] = xk; r[k] = rk;
        acc += rk * rk;
    }
    publish_sum(acc, partials, state, red, Upd{0, {}, {}}, true);
}

int blocks_for(unsigned long long n) { const unsigned long long b = (n + kThreads - 1) / kThreads; return (int)(b < (unsigned long long)kSlabBlocks ? b : kSlabBlocks); }

}  // namespace

size_t cg_slab_state_words() { return S_WORDS; }
size_t cg_slab_partial_words() { return kSlabBlocks; }

// op 0 init | 3 phase A of iteration `it` | 5 phase B | 7 after the last iteration.  r, p_old, p_new point to plane 0 of
// [nloc + 2][P] arrays (planes -1 and nloc are the halos); state: cg_slab_state_words() doubles, zeroed by the caller before op 0.
int launch_cg_slab(cudaStream_t st, int op, int gNt, int n0, int nloc, int Ny, int Nx, double rcoef, double eps, double rtol, int it,
                   int maxiter, const double *b, double *x, double *r, double *p_old, double *p_new, double *q, double *partials,
                   double *state)
{
    const unsigned long long P = (unsigned long long)Nx * Ny, N = P * nloc;
    switch (op) {
    case 0: k_init<<<blocks_for(N + 2 * P), kThreads, 0, st>>>((unsigned int)N, (unsigned int)P, b, x, r - P, p_old - P, partials, state); break;
    case 3: k_phase_a<<<blocks_for(P), kThreads, 0, st>>>(gNt, n0, nloc, Ny, Nx, rcoef, eps, rtol, it, r, p_old, p_new, q, partials, state); break;
    case 5: k_phase_b<<<blocks_for(N), kThreads, 0, st>>>((unsigned int)N, p_new, q, x, r, partials, state); break;
    case 7: k_finish<<<1, 1, 0, st>>>(state, maxiter); break;
    default: set_error("launch_cg_slab: bad op %d", op); return FOTO_ERR_ARG;
    }
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

}  // namespace foto
