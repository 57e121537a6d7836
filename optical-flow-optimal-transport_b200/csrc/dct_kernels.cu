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
// Transforms are batched dense fp64 GEMMs on the CUDA cores (584 = 8*73 and 388 = 4*97 are
// FFT-hostile; 3.5 GFLOP per solve at 388x584x4).  B200's fp64 tensor (DMMA) peak equals its
// DFMA peak, so tensor cores would not raise the roof of this kernel; the register-tiled kernel
// below keeps the fp64 pipe, not shared memory, as the limiter (6 shared-memory wavefronts
// against 32 fp64 issue cycles per k-step and warp).
#include <cmath>

#include "foto_kernels.cuh"

namespace foto {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;     // 256 threads, 4x4 outputs each

// C[b] = A[b] * B[b], row-major, A: M x K (lda), B: K x N (ldb), C: M x N (ldc); batch strides in doubles
__global__ void __launch_bounds__(256) k_dgemm_nn(int M, int N, int K, const double *__restrict__ A, int lda,
                                                   long long strideA, const double *__restrict__ B, int ldb,
                                                   long long strideB, double *__restrict__ C, int ldc, long long strideC)
{
    __shared__ __align__(16) double As[2][BK][BM + 4];      // As[k][m] (transposed on load)
    __shared__ __align__(16) double Bs[2][BK][BN];
    A += (size_t)blockIdx.z * strideA; B += (size_t)blockIdx.z * strideB; C += (size_t)blockIdx.z * strideC;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;          // thread tile: rows ty*4.., cols tx*4..
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.0;

    // loaders: A tile 64 x 16 (each thread 4 elements: row = tid/4, cols (tid%4)*4..), B tile 16 x 64
    const int ar = tid >> 2, ac = (tid & 3) * 4;
    const int br = tid >> 4, bc = (tid & 15) * 4;
    auto load_tiles = [&](int buf, int k0) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int gm = m0 + ar, gk = k0 + ac + e;
            As[buf][ac + e][ar] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.0;
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int gk = k0 + br, gn = n0 + bc + e;
            Bs[buf][br][bc + e] = (gk < K && gn < N) ? B[(size_t)gk * ldb + gn] : 0.0;
        }
    };
    load_tiles(0, 0);
    __syncthreads();
    const int nk = (K + BK - 1) / BK;
    for (int kt = 0; kt < nk; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(buf ^ 1, (kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            double a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = As[buf][k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; j++) b[j] = Bs[buf][k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int gm = m0 + ty * TM + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int gn = n0 + tx * TN + j;
            if (gn < N) C[(size_t)gm * ldc + gn] = acc[i][j];
        }
    }
}

// divide the spectrum by the eigenvalues of A: r (eps + lam_t[a] + lam_y[b] + lam_x[c])
__global__ void __launch_bounds__(256) k_spectral_divide(int Nt, int Ny, int Nx, double r, double eps,
                                                          const double *__restrict__ lam_t, const double *__restrict__ lam_y,
                                                          const double *__restrict__ lam_x, double *__restrict__ v)
{
    const unsigned int N = (unsigned int)Nt * Ny * Nx, stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += stride) {
        const unsigned int row = k / (unsigned int)Nx, x = k - row * Nx, t = row / (unsigned int)Ny, y = row - t * Ny;
        v[k] = v[k] / (r * (eps + lam_t[t] + lam_y[y] + lam_x[x]));
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
    k_dgemm_nn<<<grid, 256, 0, st>>>(M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC);
}

// phi = A^-1 F.  tabs: device tables for this grid; w0, w1: two N-double scratch volumes.
// 7 launches: x, y, t forward transforms, spectral divide, t, y, x inverse transforms.
int launch_poisson_dct(cudaStream_t st, const DctTables &tb, int Nt, int Ny, int Nx, double r, double eps,
                       const double *F, double *phi, double *w0, double *w1)
{
    const long long P = (long long)Ny * Nx;
    // forward: rows * Cx^T  (M = Nt*Ny, K = Nx),   Cy * plane (batched over t),   Ct * [Nt x P]
    gemm(st, Nt * Ny, Nx, Nx, F, Nx, 0, tb.CxT, Nx, 0, w0, Nx, 0, 1);
    gemm(st, Ny, Nx, Ny, tb.Cy, Ny, 0, w0, Nx, P, w1, Nx, P, Nt);
    gemm(st, Nt, (int)P, Nt, tb.Ct, Nt, 0, w1, (int)P, 0, w0, (int)P, 0, 1);
    k_spectral_divide<<<148 * 8, 256, 0, st>>>(Nt, Ny, Nx, r, eps, tb.lam_t, tb.lam_y, tb.lam_x, w0);
    // inverse (DCT-III = transpose)
    gemm(st, Nt, (int)P, Nt, tb.CtT, Nt, 0, w0, (int)P, 0, w1, (int)P, 0, 1);
    gemm(st, Ny, Nx, Ny, tb.CyT, Ny, 0, w1, Nx, P, w0, Nx, P, Nt);
    gemm(st, Nt * Ny, Nx, Nx, w0, Nx, 0, tb.Cx, Nx, 0, phi, Nx, 0, 1);
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

}  // namespace foto
