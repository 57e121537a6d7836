// dct_kernels.cu -- K2b: exact Poisson solve of stepA by separable DCT-II / DCT-III.
//
// The system matrix of stepA is A = r (eps I - L_t - L_y - L_x) with the 1-D Neumann Laplacians
// of operators.lap1d (operators.py:95-110), and L = C^T diag(-4 sin^2(pi k / 2n)) C with C the
// orthonormal DCT-II matrix (off-diagonals 4.5e-16).  Hence
//     phi = (Ct^T (x) Cy^T (x) Cx^T) [ (Ct (x) Cy (x) Cx) F  /  r (eps + lt_a + ly_b + lx_c) ].
// This replaces the ~670 iterations of the reference's truncated CG by six dense transforms.  It
// is NOT what the reference computes: the reference stops its CG at rtol 1e-6, and the exact
// solve moves the final flow by ~5e-7 relative (SURVEY.md parity trap #1).  It is therefore an
// opt-in back-end (FOTO_POISSON_DCT_EXACT), gated against the "tight" goldens (the reference
// with its inner CG run to rtol 1e-13), never the default.
//
// Transforms are batched dense fp64 GEMMs (584 = 8*73 and 388 = 4*97 are FFT-hostile; 3.5 GFLOP
// per solve at 388x584x4) on the fp64 tensor-core MMA (m8n8k4, SASS DMMA): see k_dgemm_nn.
#include <cmath>

#include "foto_kernels.cuh"

namespace foto {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, GT = 128, STAGES = 2;   // 4 warps, each a 32x32 output tile
constexpr int APITCH = BK + 4;       // = 4 (mod 16): the 8x4 A fragment of a half-warp hits 16 distinct bank pairs
constexpr int BPITCH = BN + 4;       // = 4 (mod 16): same for the 4x8 B fragment

// 8-byte asynchronous global -> shared copy (LDGSTS), zero-filled when !valid
__device__ __forceinline__ void cp_async8(double *dst, const double *src, bool valid)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor-core MMA (SASS: DMMA)
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C[b] = A[b] * B[b], row-major, A: M x K (lda), B: K x N (ldb), C: M x N (ldc); batch strides in doubles.
// A register-tiled CUDA-core version of this kernel (4x4 outputs per thread) ran at 17 % of the fp64
// peak with the shared-memory pipe 50-58 % busy (ncu): operand delivery, not the fp64 pipe, was the
// limit.  The m8n8k4 fp64 MMA shares each operand across the warp inside the tensor-core datapath:
// a warp tile of 32x32 needs 8 shared-memory doubles per thread for 128 FMAs per thread (the 4x4
// CUDA-core tile: 8 doubles for 16 FMAs).  Operand tiles travel global -> shared with cp.async, one
// k-tile ahead of the MMAs.
__global__ void __launch_bounds__(GT) k_dgemm_nn(int M, int N, int K, const double *__restrict__ A, int lda,
                                                  long long strideA, const double *__restrict__ B, int ldb,
                                                  long long strideB, double *__restrict__ C, int ldc, long long strideC)
{
    __shared__ __align__(16) double As[STAGES][BM][APITCH];   // As[m][k]
    __shared__ __align__(16) double Bs[STAGES][BK][BPITCH];   // Bs[k][n]
    A += (size_t)blockIdx.z * strideA; B += (size_t)blockIdx.z * strideB; C += (size_t)blockIdx.z * strideC;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;              // warp tile origin inside the block tile
    const int fr = lane >> 2, fc = lane & 3;                            // fragment row / column of this lane
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // A tile 64 x 16: 8 passes of (row = pass*8 + tid/16, k = tid%16); B tile 16 x 64: 8 passes of (row = pass*2 + tid/64, col = tid%64)
    auto prefetch = [&](int stage, int k0) {
#pragma unroll
        for (int ps = 0; ps < 8; ps++) {
            const int r = ps * 8 + (tid >> 4), c = tid & 15;
            const int gm = m0 + r, gk = k0 + c;
            const bool ok = gm < M && gk < K;
            cp_async8(&As[stage][r][c], ok ? A + (size_t)gm * lda + gk : A, ok);
        }
#pragma unroll
        for (int ps = 0; ps < 8; ps++) {
            const int r = ps * 2 + (tid >> 6), c = tid & 63;
            const int gk = k0 + r, gn = n0 + c;
            const bool ok = gk < K && gn < N;
            cp_async8(&Bs[stage][r][c], ok ? B + (size_t)gk * ldb + gn : B, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nk = (K + BK - 1) / BK;
    prefetch(0, 0);
    for (int kt = 0; kt < nk; kt++) {
        const int st = kt & 1;
        if (kt + 1 < nk) prefetch(st ^ 1, (kt + 1) * BK); else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");     // k-tile kt has landed
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[st][wm + i * 8 + fr][k4 + fc];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[st][k4 + fc][wn + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();                                          // stage st may be overwritten by the next prefetch
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int gm = m0 + wm + i * 8 + fr;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gn = n0 + wn + j * 8 + fc * 2;
            if (gn < N) C[(size_t)gm * ldc + gn] = acc[i][j][0];
            if (gn + 1 < N) C[(size_t)gm * ldc + gn + 1] = acc[i][j][1];
        }
    }
}

// Forward t-DCT, division by the eigenvalues r (eps + lam_t[a] + lam_y[b] + lam_x[c]) and inverse
// t-DCT fused into one pass: one thread per (y, x) spectral column, Nt values in registers.
template <int MAXNT>
__global__ void __launch_bounds__(256) k_t_solve(int Nt, int Ny, int Nx, double r, double eps,
                                                  const double *__restrict__ Ct, const double *__restrict__ lam_t,
                                                  const double *__restrict__ lam_y, const double *__restrict__ lam_x,
                                                  const double *__restrict__ in, double *__restrict__ out)
{
    __shared__ double sC[MAXNT * MAXNT], sl[MAXNT];
    for (int i = threadIdx.x; i < Nt * Nt; i += blockDim.x) sC[i] = Ct[i];
    for (int i = threadIdx.x; i < Nt; i += blockDim.x) sl[i] = lam_t[i];
    __syncthreads();
    const unsigned int P = (unsigned int)Ny * Nx;
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const unsigned int y = k / (unsigned int)Nx, x = k - y * Nx;
    const double lyx = lam_y[y] + lam_x[x];
    double v[MAXNT], w[MAXNT];
#pragma unroll
    for (int n = 0; n < MAXNT; n++) v[n] = n < Nt ? in[(size_t)n * P + k] : 0.0;
#pragma unroll
    for (int a = 0; a < MAXNT; a++) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < MAXNT; n++) if (n < Nt) s = fma(sC[a * Nt + n], v[n], s);
        w[a] = a < Nt ? s / (r * (eps + sl[a] + lyx)) : 0.0;
    }
#pragma unroll
    for (int n = 0; n < MAXNT; n++) {
        if (n < Nt) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < MAXNT; a++) if (a < Nt) s = fma(sC[a * Nt + n], w[a], s);
            out[(size_t)n * P + k] = s;
        }
    }
}

}  // namespace

// Orthonormal DCT-II matrix C[k][i] = s_k cos(pi k (2i+1) / (2n)) (row k = k-th basis vector), its
// transpose, and the eigenvalues 4 sin^2(pi k / 2n) of -lap1d('N').  The cosine argument is reduced
// modulo 4n in integers first, so the entries are accurate to an ulp for any n.
void dct_host_tables(int n, std::vector<double> &C, std::vector<double> &Ct, std::vector<double> &lam)
{
    C.assign((size_t)n * n, 0.0); Ct.assign((size_t)n * n, 0.0); lam.assign(n, 0.0);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < n; k++) {
        const double s = k == 0 ? std::sqrt(1.0 / n) : std::sqrt(2.0 / n);
        for (int i = 0; i < n; i++) {
            const long long m = ((long long)k * (2 * i + 1)) % (4LL * n);       // angle = pi * m / (2n), period 4n
            const double c = s * std::cos(pi * (double)m / (2.0 * n));
            C[(size_t)k * n + i] = c;
            Ct[(size_t)i * n + k] = c;
        }
        const double sn = std::sin(pi * (double)k / (2.0 * n));
        lam[k] = 4.0 * sn * sn;
    }
}

static void gemm(cudaStream_t st, int M, int N, int K, const double *A, int lda, long long sA, const double *B, int ldb,
                 long long sB, double *C, int ldc, long long sC, int batch)
{
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
    k_dgemm_nn<<<grid, GT, 0, st>>>(M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC);
}

static int t_solve(cudaStream_t st, const DctTables &tb, int Nt, int Ny, int Nx, const double *lam_y, double r, double eps,
                   const double *in, double *out)
{
    const int blocks = (int)(((long long)Ny * Nx + 255) / 256);
#define FOTO_T_SOLVE(M) k_t_solve<M><<<blocks, 256, 0, st>>>(Nt, Ny, Nx, r, eps, tb.Ct, tb.lam_t, lam_y, tb.lam_x, in, out)
    if (Nt <= 4) FOTO_T_SOLVE(4);
    else if (Nt <= 8) FOTO_T_SOLVE(8);
    else if (Nt <= 16) FOTO_T_SOLVE(16);
    else if (Nt <= 32) FOTO_T_SOLVE(32);
    else if (Nt <= 64) FOTO_T_SOLVE(64);
    else { set_error("dct_exact supports Nt <= 64"); return FOTO_ERR_ARG; }
#undef FOTO_T_SOLVE
    return FOTO_OK;
}

// x and y transforms of `nplanes` planes (forward: DCT-II, inverse: DCT-III); tmp: nplanes*Ny*Nx doubles
int launch_dct_xy(cudaStream_t st, const DctTables &tb, int nplanes, int Ny, int Nx, const double *in, double *out,
                  double *tmp, int inverse)
{
    const long long P = (long long)Ny * Nx;
    if (!inverse) {
        gemm(st, nplanes * Ny, Nx, Nx, in, Nx, 0, tb.CxT, Nx, 0, tmp, Nx, 0, 1);       // rows * Cx^T
        gemm(st, Ny, Nx, Ny, tb.Cy, Ny, 0, tmp, Nx, P, out, Nx, P, nplanes);            // Cy * plane
    } else {
        gemm(st, Ny, Nx, Ny, tb.CyT, Ny, 0, in, Nx, P, tmp, Nx, P, nplanes);
        gemm(st, nplanes * Ny, Nx, Nx, tmp, Nx, 0, tb.Cx, Nx, 0, out, Nx, 0, 1);
    }
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

// t transform, division by the eigenvalues and inverse t transform on a [Nt][ny_loc][Nx] block whose rows
// are the global rows y_off .. y_off + ny_loc - 1 (time-slab mode after the all-to-all transpose)
int launch_dct_t_solve(cudaStream_t st, const DctTables &tb, int Nt, int ny_loc, int Nx, int y_off, double r, double eps,
                       const double *in, double *out)
{
    FOTO_TRY(t_solve(st, tb, Nt, ny_loc, Nx, tb.lam_y + y_off, r, eps, in, out));
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

// phi = A^-1 F.  tb: device tables for this grid; w0, w1: two N-double scratch volumes.
// 5 launches: x, y forward transforms; fused t-transform / divide / inverse t; y, x inverse transforms.
int launch_poisson_dct(cudaStream_t st, const DctTables &tb, int Nt, int Ny, int Nx, double r, double eps,
                       const double *F, double *phi, double *w0, double *w1)
{
    FOTO_TRY(launch_dct_xy(st, tb, Nt, Ny, Nx, F, w1, w0, 0));
    FOTO_TRY(t_solve(st, tb, Nt, Ny, Nx, tb.lam_y, r, eps, w1, w0));
    FOTO_TRY(launch_dct_xy(st, tb, Nt, Ny, Nx, w0, phi, w1, 1));
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

}  // namespace foto
