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
// per solve at 388x584x4) on the fp64 tensor-core MMA (m8n8k4, SASS DMMA): see k_dgemm_nn.  When Nx and Ny are
// multiples of 4 the even / odd symmetry of the DCT matrix halves the flops (k_fold / k_unfold, launch_dct_xy).
#include <cmath>

#include "foto_kernels.cuh"

namespace foto {

namespace {

constexpr int BK = 16;
constexpr int APITCH = BK + 4;       // = 4 (mod 16): the 8x4 A fragment of a half-warp hits 16 distinct bank pairs

// 8- / 16-byte asynchronous global -> shared copy (LDGSTS), zero-filled when !valid
__device__ __forceinline__ void cp_async8(double *dst, const double *src, bool valid)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(double *dst, const double *src, bool valid)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// D(8x8) += A(8x4, row) * B(4x8, col), fp64 tensor-core MMA (SASS: DMMA)
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct GemmBatch {                    // two-level batch: blockIdx.z = z1 * nb0 + z0, offsets in doubles
    int nb0;
    long long a0, a1, b0, b1, c0, c1;
};

// C[z] = A[z] * B[z], row-major, A: M x K (lda), B: K x N (ldb), C: M x N (ldc).
// A register-tiled CUDA-core version of this kernel (4x4 outputs per thread) ran at 17 % of the fp64
// peak with the shared-memory pipe 50-58 % busy (ncu): operand delivery, not the fp64 pipe, was the
// limit.  The m8n8k4 fp64 MMA shares each operand across the warp inside the tensor-core datapath:
// a warp tile of 32x32 needs 8 shared-memory doubles per thread for 128 FMAs per thread (the 4x4
// CUDA-core tile: 8 doubles for 16 FMAs).  Operand tiles travel global -> shared with cp.async through a
// STAGES-deep ring; VEC: 16-byte copies (all leading dimensions, K, N and the base offsets even).
// Block tile BM x BN, (BM / 32) x (BN / 32) warps of 32 x 32 each.
template <int BM, int BN, int STAGES, bool VEC>
__global__ void __launch_bounds__((BM / 32) * (BN / 32) * 32) k_dgemm_nn(int M, int N, int K, const double *__restrict__ A, int lda,
                                                                        const double *__restrict__ B, int ldb, double *__restrict__ C,
                                                                        int ldc, GemmBatch gb)
{
    constexpr int GT = (BM / 32) * (BN / 32) * 32, BPITCH = BN + 4;       // = 4 (mod 16): same for the 4x8 B fragment
    extern __shared__ __align__(16) double dsm[];
    double (*As)[BM][APITCH] = reinterpret_cast<double (*)[BM][APITCH]>(dsm);                          // As[stage][m][k]
    double (*Bs)[BK][BPITCH] = reinterpret_cast<double (*)[BK][BPITCH]>(dsm + STAGES * BM * APITCH);   // Bs[stage][k][n]
    {
        const int z0 = blockIdx.z % gb.nb0, z1 = blockIdx.z / gb.nb0;
        A += z0 * gb.a0 + z1 * gb.a1; B += z0 * gb.b0 + z1 * gb.b1; C += z0 * gb.c0 + z1 * gb.c1;
    }
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp / (BN / 32)) * 32, wn = (warp % (BN / 32)) * 32;  // warp tile origin inside the block tile
    const int fr = lane >> 2, fc = lane & 3;                            // fragment row / column of this lane
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    auto prefetch = [&](int stage, int k0) {
        if (VEC) {
            // A tile BM x 16: 8 two-double chunks per row; B tile 16 x BN: BN / 2 chunks per row
#pragma unroll
            for (int ps = 0; ps < BM * 8 / GT; ps++) {
                const int idx = ps * GT + tid, r = idx >> 3, c = (idx & 7) * 2;
                const int gm = m0 + r, gk = k0 + c;
                const bool ok = gm < M && gk < K;
                cp_async16(&As[stage][r][c], ok ? A + (size_t)gm * lda + gk : A, ok);
            }
#pragma unroll
            for (int ps = 0; ps < BK * (BN / 2) / GT; ps++) {
                const int idx = ps * GT + tid, r = idx / (BN / 2), c = (idx % (BN / 2)) * 2;
                const int gk = k0 + r, gn = n0 + c;
                const bool ok = gk < K && gn < N;
                cp_async16(&Bs[stage][r][c], ok ? B + (size_t)gk * ldb + gn : B, ok);
            }
        } else {
#pragma unroll
            for (int ps = 0; ps < BM * BK / GT; ps++) {
                const int idx = ps * GT + tid, r = idx / BK, c = idx % BK;
                const int gm = m0 + r, gk = k0 + c;
                const bool ok = gm < M && gk < K;
                cp_async8(&As[stage][r][c], ok ? A + (size_t)gm * lda + gk : A, ok);
            }
#pragma unroll
            for (int ps = 0; ps < BK * BN / GT; ps++) {
                const int idx = ps * GT + tid, r = idx / BN, c = idx % BN;
                const int gk = k0 + r, gn = n0 + c;
                const bool ok = gk < K && gn < N;
                cp_async8(&Bs[stage][r][c], ok ? B + (size_t)gk * ldb + gn : B, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nk = (K + BK - 1) / BK;
#pragma unroll
    for (int s0 = 0; s0 < STAGES - 1; s0++) {
        if (s0 < nk) prefetch(s0, s0 * BK); else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = 0; kt < nk; kt++) {
        const int st = kt % STAGES;
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");     // k-tile kt has landed
        __syncthreads();                                                          // ... and stage (kt-1) % STAGES is free
        if (kt + STAGES - 1 < nk) prefetch((kt + STAGES - 1) % STAGES, (kt + STAGES - 1) * BK);
        else asm volatile("cp.async.commit_group;" ::: "memory");
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

// Even / odd folding of a DCT along the middle axis of an [outer][n][inner] array (n even, h = n / 2).  The DCT-II
// matrix satisfies C[k][n-1-i] = (-1)^k C[k][i], so
//   forward:  X[2j]   = sum_{i<h} C[2j][i]   (x[i] + x[n-1-i]),   X[2j+1] = sum_{i<h} C[2j+1][i] (x[i] - x[n-1-i])
//   inverse:  x[i] = E[i] + O[i],  x[n-1-i] = E[i] - O[i],  E = sum_j C[2j][i] X[2j],  O = sum_j C[2j+1][i] X[2j+1]
// i.e. two h x h transforms instead of one n x n: half the flops.  fold: out[b][outer][h][inner] (b = 0 sums, 1 differences)
__global__ void __launch_bounds__(256) k_fold(size_t total, int n, int inner, const double *__restrict__ in, double *__restrict__ out, size_t bstride)
{
    const int h = n >> 1;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const size_t i = k % inner, oh = k / inner, o = oh / h;
        const int j = (int)(oh % h);
        const double a = in[(o * n + j) * inner + i], c = in[(o * n + (n - 1 - j)) * inner + i];
        out[k] = a + c;
        out[bstride + k] = a - c;
    }
}
__global__ void __launch_bounds__(256) k_unfold(size_t total, int n, int inner, const double *__restrict__ in, size_t bstride, double *__restrict__ out)
{
    const int h = n >> 1;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const size_t i = k % inner, oh = k / inner, o = oh / h;
        const int j = (int)(oh % h);
        const double e = in[k], d = in[bstride + k];
        out[(o * n + j) * inner + i] = e + d;
        out[(o * n + (n - 1 - j)) * inner + i] = e - d;
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

// Folded tables of one axis (n even): E[b][j][i] = C[2j+b][i], i < n/2; ET[b][i][j] = E[b][j][i]; lam_p = (lam[0], lam[2], ...,
// lam[1], lam[3], ...)
void dct_host_folded(int n, const std::vector<double> &C, const std::vector<double> &lam, std::vector<double> &E,
                     std::vector<double> &ET, std::vector<double> &lam_p)
{
    const int h = n / 2;
    E.assign((size_t)2 * h * h, 0.0); ET.assign((size_t)2 * h * h, 0.0); lam_p.assign(n, 0.0);
    for (int b = 0; b < 2; b++)
        for (int j = 0; j < h; j++) {
            lam_p[b * h + j] = lam[2 * j + b];
            for (int i = 0; i < h; i++) {
                const double c = C[(size_t)(2 * j + b) * n + i];
                E[((size_t)b * h + j) * h + i] = c;
                ET[((size_t)b * h + i) * h + j] = c;
            }
        }
}

// Second level (n a multiple of 8, q = n / 4): the even half is itself a DCT-II of length h = n / 2 on e1[i] = x[i] + x[n-1-i]
// (C_n[2j][h-1-i] = (-1)^j C_n[2j][i]), so it folds once more: three transforms q x q, q x q, h x h = 3/8 of the dense flops, and
// still one pass over the data each way.  Blocks of the folded array: [e2 | o2 | o1] of q, q and h rows of `inner` words per
// outer index; spectrum order: frequencies 4m, then 4m+2, then 2j+1.
__global__ void __launch_bounds__(256) k_fold2(size_t total, int n, int inner, const double *__restrict__ in, double *__restrict__ out,
                                                size_t outer)
{
    const int h = n >> 1, q = n >> 2;
    const size_t o2off = outer * q * inner, o1off = 2 * o2off;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const size_t i = k % inner, oh = k / inner, o = oh / h;
        const int j = (int)(oh % h);
        const double a = in[(o * n + j) * inner + i], b = in[(o * n + (n - 1 - j)) * inner + i];
        out[o1off + k] = a - b;
        if (j < q) {
            const double c = in[(o * n + (h - 1 - j)) * inner + i], d = in[(o * n + (h + j)) * inner + i];
            const double e1a = a + b, e1b = c + d;
            const size_t kk = (o * q + j) * inner + i;
            out[kk] = e1a + e1b;
            out[o2off + kk] = e1a - e1b;
        }
    }
}
__global__ void __launch_bounds__(256) k_unfold2(size_t total, int n, int inner, const double *__restrict__ in, double *__restrict__ out,
                                                  size_t outer)
{
    const int h = n >> 1, q = n >> 2;
    const size_t o2off = outer * q * inner, o1off = 2 * o2off;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {   // total = outer q inner
        const size_t i = k % inner, oq = k / inner, o = oq / q;
        const int j = (int)(oq % q);
        const double E2 = in[k], O2 = in[o2off + k];
        const double E1a = E2 + O2, E1b = E2 - O2;                      // E1[j], E1[h-1-j]
        const double O1a = in[o1off + (o * h + j) * inner + i], O1b = in[o1off + (o * h + (h - 1 - j)) * inner + i];
        out[(o * n + j) * inner + i] = E1a + O1a;
        out[(o * n + (n - 1 - j)) * inner + i] = E1a - O1a;
        out[(o * n + (h - 1 - j)) * inner + i] = E1b + O1b;
        out[(o * n + (h + j)) * inner + i] = E1b - O1b;
    }
}

// Second-level tables of one axis (n a multiple of 8, q = n / 4): E2[c][m][i] = C[4m + 2c][i], i < q; E2T its transposes;
// lam_p2 = (lam[4m] | lam[4m+2] | lam[2j+1])
void dct_host_folded2(int n, const std::vector<double> &C, const std::vector<double> &lam, std::vector<double> &E2,
                      std::vector<double> &E2T, std::vector<double> &lam_p2)
{
    const int h = n / 2, q = n / 4;
    E2.assign((size_t)2 * q * q, 0.0); E2T.assign((size_t)2 * q * q, 0.0); lam_p2.assign(n, 0.0);
    for (int c = 0; c < 2; c++)
        for (int m = 0; m < q; m++) {
            lam_p2[c * q + m] = lam[4 * m + 2 * c];
            for (int i = 0; i < q; i++) {
                const double v = C[(size_t)(4 * m + 2 * c) * n + i];
                E2[((size_t)c * q + m) * q + i] = v;
                E2T[((size_t)c * q + i) * q + m] = v;
            }
        }
    for (int j = 0; j < h; j++) lam_p2[h + j] = lam[2 * j + 1];
}

template <int BM, int BN, int STAGES>
struct DGemm {
    static constexpr int smem = STAGES * (BM * APITCH + BK * (BN + 4)) * (int)sizeof(double);
    static constexpr int threads = (BM / 32) * (BN / 32) * 32;
    static int launch(cudaStream_t st, bool vec, int M, int N, int K, const double *A, int lda, const double *B, int ldb, double *C,
                      int ldc, const GemmBatch &gb, int batch)
    {
        static bool prepared[64] = {};
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (dev < 64 && !prepared[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(k_dgemm_nn<BM, BN, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CUDA_TRY(cudaFuncSetAttribute(k_dgemm_nn<BM, BN, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            prepared[dev] = true;
        }
        dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
        if (vec) k_dgemm_nn<BM, BN, STAGES, true><<<grid, threads, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, gb);
        else k_dgemm_nn<BM, BN, STAGES, false><<<grid, threads, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, gb);
        return FOTO_OK;
    }
};
// 64 x 64 tiles, 4 warps, 3 stages, 37 KB: 31 TFLOP/s = 0.84 of the measured DMMA peak at 1080x1920x16.  128 x 128 tiles with 16
// warps (a quarter of the L2 -> shared traffic per flop) were measured slower (28 TFLOP/s; half the speed on one Middlebury pair).
using DGemm64 = DGemm<64, 64, 3>;

static bool aligned16(const void *p) { return ((size_t)p & 15) == 0; }

static int gemm(cudaStream_t st, int M, int N, int K, const double *A, int lda, const double *B, int ldb, double *C, int ldc,
                const GemmBatch &gb, int batch)
{
    const bool even = !((lda | ldb | K | N) & 1) && !((gb.a0 | gb.a1 | gb.b0 | gb.b1) & 1) && aligned16(A) && aligned16(B);
    return DGemm64::launch(st, even, M, N, K, A, lda, B, ldb, C, ldc, gb, batch);
}

static int fold_blocks(size_t total) { const size_t b = (total + 255) / 256; return (int)(b < 148 * 16 ? b : 148 * 16); }

static int t_solve(cudaStream_t st, const DctTables &tb, int Nt, int Ny, int Nx, const double *lam_y, double r, double eps,
                   const double *in, double *out)
{
    const int blocks = (int)(((long long)Ny * Nx + 255) / 256);
#define FOTO_T_SOLVE(M) k_t_solve<M><<<blocks, 256, 0, st>>>(Nt, Ny, Nx, r, eps, tb.Ct, tb.lam_t, lam_y, tb.split ? tb.lam_xp : tb.lam_x, in, out)
    if (Nt <= 4) FOTO_T_SOLVE(4);
    else if (Nt <= 8) FOTO_T_SOLVE(8);
    else if (Nt <= 16) FOTO_T_SOLVE(16);
    else if (Nt <= 32) FOTO_T_SOLVE(32);
    else if (Nt <= 64) FOTO_T_SOLVE(64);
    else { set_error("dct_exact supports Nt <= 64"); return FOTO_ERR_ARG; }
#undef FOTO_T_SOLVE
    return FOTO_OK;
}

// x and y transforms of `nplanes` planes (forward: DCT-II, inverse: DCT-III); tmp: nplanes*Ny*Nx doubles, in != out != tmp.
// tb.split (Nx and Ny multiples of 4): even / odd folded transforms, half the flops (3/8 along an axis whose length is a
// multiple of 8: tb.lx / tb.ly = 2); the spectrum then lives in a permuted order along x and y (tb.lam_xp / lam_yp are permuted
// alike; nothing but the pointwise t solve ever looks at the spectrum).
int launch_dct_xy(cudaStream_t st, const DctTables &tb, int nplanes, int Ny, int Nx, const double *in, double *out,
                  double *tmp, int inverse)
{
    const long long P = (long long)Ny * Nx;
    const GemmBatch one = {1, 0, 0, 0, 0, 0, 0};
    if (!tb.split) {
        if (!inverse) {
            FOTO_TRY(gemm(st, nplanes * Ny, Nx, Nx, in, Nx, tb.CxT, Nx, tmp, Nx, one, 1));                     // rows * Cx^T
            FOTO_TRY(gemm(st, Ny, Nx, Ny, tb.Cy, Ny, tmp, Nx, out, Nx, GemmBatch{1, 0, 0, 0, P, 0, P}, nplanes));   // Cy * plane
        } else {
            FOTO_TRY(gemm(st, Ny, Nx, Ny, tb.CyT, Ny, in, Nx, tmp, Nx, GemmBatch{1, 0, 0, 0, P, 0, P}, nplanes));
            FOTO_TRY(gemm(st, nplanes * Ny, Nx, Nx, tmp, Nx, tb.Cx, Nx, out, Nx, one, 1));
        }
        CUDA_TRY(cudaGetLastError());
        return FOTO_OK;
    }
    const int hx = Nx / 2, hy = Ny / 2, qx = Nx / 4, qy = Ny / 4, R = nplanes * Ny;
    const size_t half = (size_t)nplanes * P / 2;          // elements of one folded half-volume
    const size_t quarter = half / 2;
    const long long hyNx = (long long)hy * Nx, qyNx = (long long)qy * Nx;
    // forward x: rows of `src` -> spectrum columns of `dst` (ld Nx); the folded rows live in tmp
    auto fwd_x = [&](const double *src, double *dst) -> int {
        if (tb.lx == 2) {
            k_fold2<<<fold_blocks(half), 256, 0, st>>>(half, Nx, 1, src, tmp, (size_t)R);           // tmp = [e2 | o2 | o1]: R x qx, R x qx, R x hx
            FOTO_TRY(gemm(st, R, qx, qx, tmp, qx, tb.E2xT, qx, dst, Nx, GemmBatch{2, (long long)R * qx, 0, (long long)qx * qx, 0, qx, 0}, 2));
            FOTO_TRY(gemm(st, R, hx, hx, tmp + 2 * (size_t)R * qx, hx, tb.ExT + (size_t)hx * hx, hx, dst + hx, Nx, one, 1));
        } else {
            k_fold<<<fold_blocks(half), 256, 0, st>>>(half, Nx, 1, src, tmp, half);                 // tmp = [b][R][hx]
            FOTO_TRY(gemm(st, R, hx, hx, tmp, hx, tb.ExT, hx, dst, Nx, GemmBatch{2, (long long)R * hx, 0, (long long)hx * hx, 0, hx, 0}, 2));
        }
        return FOTO_OK;
    };
    auto fwd_y = [&](const double *src, double *dst) -> int {
        if (tb.ly == 2) {
            k_fold2<<<fold_blocks(half), 256, 0, st>>>(half, Ny, Nx, src, tmp, (size_t)nplanes);    // tmp = [e2 | o2 | o1][plane][rows][Nx]
            FOTO_TRY(gemm(st, qy, Nx, qy, tb.E2y, qy, tmp, Nx, dst, Nx,
                          GemmBatch{2, (long long)qy * qy, 0, (long long)nplanes * qyNx, qyNx, qyNx, P}, 2 * nplanes));
            FOTO_TRY(gemm(st, hy, Nx, hy, tb.Ey + (size_t)hy * hy, hy, tmp + 2 * (size_t)nplanes * qyNx, Nx, dst + hyNx, Nx,
                          GemmBatch{1, 0, 0, 0, hyNx, 0, P}, nplanes));
        } else {
            k_fold<<<fold_blocks(half), 256, 0, st>>>(half, Ny, Nx, src, tmp, half);                // tmp = [b][plane][hy][Nx]
            FOTO_TRY(gemm(st, hy, Nx, hy, tb.Ey, hy, tmp, Nx, dst, Nx,
                          GemmBatch{2, (long long)hy * hy, 0, (long long)nplanes * hyNx, hyNx, hyNx, P}, 2 * nplanes));
        }
        return FOTO_OK;
    };
    auto inv_y = [&](const double *src, double *dst) -> int {
        if (tb.ly == 2) {
            FOTO_TRY(gemm(st, qy, Nx, qy, tb.E2yT, qy, src, Nx, tmp, Nx,
                          GemmBatch{2, (long long)qy * qy, 0, qyNx, P, (long long)nplanes * qyNx, qyNx}, 2 * nplanes));
            FOTO_TRY(gemm(st, hy, Nx, hy, tb.EyT + (size_t)hy * hy, hy, src + hyNx, Nx, tmp + 2 * (size_t)nplanes * qyNx, Nx,
                          GemmBatch{1, 0, 0, 0, P, 0, hyNx}, nplanes));
            k_unfold2<<<fold_blocks(quarter), 256, 0, st>>>(quarter, Ny, Nx, tmp, dst, (size_t)nplanes);
        } else {
            FOTO_TRY(gemm(st, hy, Nx, hy, tb.EyT, hy, src, Nx, tmp, Nx,
                          GemmBatch{2, (long long)hy * hy, 0, hyNx, P, (long long)nplanes * hyNx, hyNx}, 2 * nplanes));
            k_unfold<<<fold_blocks(half), 256, 0, st>>>(half, Ny, Nx, tmp, half, dst);
        }
        return FOTO_OK;
    };
    auto inv_x = [&](const double *src, double *dst) -> int {
        if (tb.lx == 2) {
            FOTO_TRY(gemm(st, R, qx, qx, src, Nx, tb.E2x, qx, tmp, qx, GemmBatch{2, qx, 0, (long long)qx * qx, 0, (long long)R * qx, 0}, 2));
            FOTO_TRY(gemm(st, R, hx, hx, src + hx, Nx, tb.Ex + (size_t)hx * hx, hx, tmp + 2 * (size_t)R * qx, hx, one, 1));
            k_unfold2<<<fold_blocks(quarter), 256, 0, st>>>(quarter, Nx, 1, tmp, dst, (size_t)R);
        } else {
            FOTO_TRY(gemm(st, R, hx, hx, src, Nx, tb.Ex, hx, tmp, hx, GemmBatch{2, hx, 0, (long long)hx * hx, 0, (long long)R * hx, 0}, 2));
            k_unfold<<<fold_blocks(half), 256, 0, st>>>(half, Nx, 1, tmp, half, dst);
        }
        return FOTO_OK;
    };
    if (!inverse) {
        FOTO_TRY(fwd_x(in, out));
        FOTO_TRY(fwd_y(out, out));                        // the y fold reads `out` into tmp before the GEMMs overwrite it
    } else {
        FOTO_TRY(inv_y(in, out));
        FOTO_TRY(inv_x(out, out));                        // the x GEMMs read `out` into tmp before the unfold overwrites it
    }
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

// t transform, division by the eigenvalues and inverse t transform on a [Nt][ny_loc][Nx] block whose rows
// are the global rows y_off .. y_off + ny_loc - 1 (time-slab mode after the all-to-all transpose)
int launch_dct_t_solve(cudaStream_t st, const DctTables &tb, int Nt, int ny_loc, int Nx, int y_off, double r, double eps,
                       const double *in, double *out)
{
    FOTO_TRY(t_solve(st, tb, Nt, ny_loc, Nx, (tb.split ? tb.lam_yp : tb.lam_y) + y_off, r, eps, in, out));
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

// phi = A^-1 F.  tb: device tables for this grid; w0, w1: two N-double scratch volumes.
// 5 launches: x, y forward transforms; fused t-transform / divide / inverse t; y, x inverse transforms.
int launch_poisson_dct(cudaStream_t st, const DctTables &tb, int Nt, int Ny, int Nx, double r, double eps,
                       const double *F, double *phi, double *w0, double *w1)
{
    FOTO_TRY(launch_dct_xy(st, tb, Nt, Ny, Nx, F, w1, w0, 0));
    FOTO_TRY(t_solve(st, tb, Nt, Ny, Nx, tb.split ? tb.lam_yp : tb.lam_y, r, eps, w1, w0));
    FOTO_TRY(launch_dct_xy(st, tb, Nt, Ny, Nx, w0, phi, w1, 1));
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

}  // namespace foto
