// gn_dct.cu -- K6 with a spectral preconditioner on the tensor cores: the Gennert-Negahdaripour system
// (classical.py:102-126) solved by fp64 conjugate gradients preconditioned with the exact inverse of its
// constant-coefficient part, applied in TF32.
//
//     A = diag(alpha, alpha, lambda) (x) (-Lap_Neumann) + g g^T (pointwise),  g = (fx, fy, -f2)
//     M = diag(alpha, alpha, lambda) (x) (-Lap_Neumann) + mean_pixels(g g^T)
// -Lap_Neumann is diagonalised by the 2-D orthonormal DCT-II (eigenvalues ly[a] + lx[b]), so
//     M^-1 r = C2^T [ (diag(alpha mu, alpha mu, lambda mu) + Gbar)^-1 (C2 r) ]   -- a 3x3 solve per frequency.
// Jacobi-PCG needs 1 270 iterations at 388x584 (kappa ~ 1.8e4, all of it from the Laplacian); with M it needs ~90
// (numpy prototype, same solution to 9e-13), and the iteration count does not change when M^-1 is applied in
// fp32 or TF32 instead of fp64 (69 / 70 / 71 iterations at 97x146): the preconditioner only has to be a fixed,
// nearly symmetric approximation, the Krylov recurrences and the residual stay in fp64.  So the four dense
// transforms of one application (2.6 GFLOP at 388x584, the cost that kept this out of round 1: 0.15 ms on the fp64
// DMMA path) run as TF32 tensor-core GEMMs (mma.sync m16n8k8, fp32 accumulation) on fp32 copies of r.
//
// The reference factorises A (SuperLU), so the solve is not tied to a Krylov sequence: any iteration converged to
// ||r|| <= 1e-13 ||b|| is a valid stand-in (tests: 1e-9 against the direct solve).
//
// One PCG iteration = 7 stream-ordered launches (single-reduction arrangement, as gn_fused.cu):
//     X, Y forward GEMMs, per-frequency 3x3 solve, Y, X inverse GEMMs      u = M^-1 r   (fp32)
//     k_stencil_dots   w = A u, partial sums of gamma = r.u, delta = w.u, rho = r.r
//     k_update         every block sums the partials in the same order; stop if rho <= rtol^2 rho_0; else
//                      beta = gamma/gamma_old, alpha = gamma/(delta - beta gamma/alpha_old),
//                      p = u + beta p, s = w + beta s, x += alpha p, r -= alpha s, r32 = float(r)
// The host enqueues iterations in chunks and reads the `done` flag one chunk behind (kernels of iterations past
// convergence return at once), so there is no host round trip per iteration.
#include "foto_kernels.cuh"

namespace foto {

namespace {

struct GnDctState {             // device
    double gam_old[2], alpha_old[2], d_old[2];  // double buffered by iteration parity
    double stop2;
    int done, iters, info, pad;
};

constexpr int BK = 32, APITCH = BK + 4;                   // A fragment loads hit bank 4g + t: conflict free

__device__ __forceinline__ void cp_async16(float *dst, const float *src, bool valid)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned int (&a)[4], const unsigned int (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// C[b] = A[b] * B[b], fp32 row-major, TF32 tensor-core MMA with fp32 accumulation.  All leading dimensions, K and the
// base pointers are multiples of 4 floats (the buffers are padded), so tiles travel with 16-byte cp.async; rows / columns
// beyond M / N are zero-filled and not stored.  Block tile BM x BN, warp tile WM x WN (m16n8 fragments), STAGES-deep
// cp.async ring.  The transforms of one image are small (0.8 GFLOP) and their operands live in L2: with 64 x 64 tiles
// the kernel re-reads 51 MB per GEMM and runs at the L2 bandwidth (15 us, 3.4 TB/s, ncu); 128-wide tiles halve that.
template <int BM, int BN, int WM, int WN, int STAGES>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) k_sgemm_tf32(int M, int N, int K, const float *__restrict__ A, int lda,
                                                                          long long strideA, const float *__restrict__ B, int ldb,
                                                                          long long strideB, float *__restrict__ C, int ldc,
                                                                          long long strideC, const int *skip)
{
    constexpr int GT = (BM / WM) * (BN / WN) * 32, BPITCH = BN + 8;      // B fragment loads hit bank 8t + g: conflict free
    constexpr int MI = WM / 16, NJ = WN / 8;
    extern __shared__ __align__(16) float gsm[];
    float (*As)[BM][APITCH] = reinterpret_cast<float (*)[BM][APITCH]>(gsm);                           // As[stage][m][k]
    float (*Bs)[BK][BPITCH] = reinterpret_cast<float (*)[BK][BPITCH]>(gsm + STAGES * BM * APITCH);    // Bs[stage][k][n]
    if (skip && *skip) return;
    A += (size_t)blockIdx.z * strideA; B += (size_t)blockIdx.z * strideB; C += (size_t)blockIdx.z * strideC;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp / (BN / WN)) * WM, wn = (warp % (BN / WN)) * WN;
    const int g = lane >> 2, t = lane & 3;
    float acc[MI][NJ][4];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[i][j][e] = 0.f;

    // A tile BM x 32 floats: 8 16-byte chunks per row; B tile 32 x BN floats: BN/4 chunks per row
    auto prefetch = [&](int stage, int k0) {
#pragma unroll
        for (int ps = 0; ps < BM * 8 / GT; ps++) {
            const int idx = ps * GT + tid, r = idx >> 3, c = (idx & 7) * 4;
            const int gm = m0 + r, gk = k0 + c;
            const bool ok = gm < M && gk < K;
            cp_async16(&As[stage][r][c], ok ? A + (size_t)gm * lda + gk : A, ok);
        }
#pragma unroll
        for (int ps = 0; ps < BK * (BN / 4) / GT; ps++) {
            const int idx = ps * GT + tid, r = idx / (BN / 4), c = (idx % (BN / 4)) * 4;
            const int gk = k0 + r, gn = n0 + c;
            const bool ok = gk < K && gn < N;
            cp_async16(&Bs[stage][r][c], ok ? B + (size_t)gk * ldb + gn : B, ok);
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
        for (int k8 = 0; k8 < BK; k8 += 8) {
            unsigned int a[MI][4], b[NJ][2];
            // A fragment (16 x 8, row major): a0 (g, t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4)
#pragma unroll
            for (int i = 0; i < MI; i++) {
                const float *p = &As[st][wm + i * 16 + g][k8 + t];
                a[i][0] = __float_as_uint(p[0]); a[i][1] = __float_as_uint(p[8 * APITCH]);
                a[i][2] = __float_as_uint(p[4]); a[i][3] = __float_as_uint(p[8 * APITCH + 4]);
            }
            // B fragment (8 x 8, column major): b0 (k = t, n = g), b1 (k = t+4, n = g)
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const float *p = &Bs[st][k8 + t][wn + j * 8 + g];
                b[j][0] = __float_as_uint(p[0]); b[j][1] = __float_as_uint(p[4 * BPITCH]);
            }
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) mma_tf32(acc[i][j], a[i], b[j]);
        }
    }
    // C fragment: c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            const int gn = n0 + wn + j * 8 + 2 * t;
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const int gm = m0 + wm + i * 16 + g + hh * 8;
                if (gm < M && gn < N)                    // N is even (multiple of 4): both columns are in range
                    *reinterpret_cast<float2 *>(C + (size_t)gm * ldc + gn) = make_float2(acc[i][j][2 * hh], acc[i][j][2 * hh + 1]);
            }
        }
}

template <int BM, int BN, int WM, int WN, int STAGES>
struct GemmCfg {
    static constexpr int smem = STAGES * (BM * APITCH + BK * (BN + 8)) * (int)sizeof(float);
    static constexpr int threads = (BM / WM) * (BN / WN) * 32;
    static void launch(cudaStream_t st, int M, int N, int K, const float *A, int lda, long long sA, const float *B, int ldb, long long sB,
                       float *C, int ldc, long long sC, int batch, const int *skip)
    {
        dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
        k_sgemm_tf32<BM, BN, WM, WN, STAGES><<<grid, threads, smem, st>>>(M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, skip);
    }
    static cudaError_t prepare() { return cudaFuncSetAttribute(k_sgemm_tf32<BM, BN, WM, WN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }
};
using GemmSmall = GemmCfg<64, 64, 32, 32, 4>;       // 128 threads, 74 KB
using GemmWide = GemmCfg<128, 128, 64, 32, 3>;      // 256 threads, 107 KB
using GemmTall = GemmCfg<128, 64, 32, 32, 3>;       // 256 threads, 81 KB

// fp64 n x n matrix -> fp32 np x np, zero padded
__global__ void k_pad_matrix(int n, int np, const double *__restrict__ in, float *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np * np) return;
    const int r = k / np, c = k - r * np;
    out[k] = (r < n && c < n) ? (float)in[(size_t)r * n + c] : 0.f;
}

// partial sums of the six entries of g g^T, g = (fx, fy, -f2)
__global__ void __launch_bounds__(256) k_gbar_partial(unsigned int P, const double *__restrict__ fx, const double *__restrict__ fy,
                                                       const double *__restrict__ f2, double *__restrict__ partials)
{
    __shared__ double red[32 * 4];
    double a[4] = {0, 0, 0, 0}, b[2] = {0, 0};
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < P; k += stride) {
        const double x = fx[k], y = fy[k], z = -f2[k];
        a[0] += x * x; a[1] += x * y; a[2] += x * z; a[3] += y * y; b[0] += y * z; b[1] += z * z;
    }
    block_sum<4>(a, red);
    block_sum<2>(b, red);
    if (threadIdx.x == 0) {
        double *p = partials + 6 * blockIdx.x;
        p[0] = a[0]; p[1] = a[1]; p[2] = a[2]; p[3] = a[3]; p[4] = b[0]; p[5] = b[1];
    }
}
// Gbar (xx, xy, xz, yy, yz, zz) = sums / P; also resets the PCG state
__global__ void __launch_bounds__(256) k_gbar_final(const double *__restrict__ partials, int blocks, double invP, double *__restrict__ gbar,
                                                     GnDctState *st)
{
    __shared__ double red[32 * 4];
    double a[4] = {0, 0, 0, 0}, b[2] = {0, 0};
    for (int i = threadIdx.x; i < blocks; i += blockDim.x) {
        const double *p = partials + 6 * i;
        a[0] += p[0]; a[1] += p[1]; a[2] += p[2]; a[3] += p[3]; b[0] += p[4]; b[1] += p[5];
    }
    block_sum<4>(a, red);
    block_sum<2>(b, red);
    if (threadIdx.x == 0) {
        gbar[0] = a[0] * invP; gbar[1] = a[1] * invP; gbar[2] = a[2] * invP; gbar[3] = a[3] * invP; gbar[4] = b[0] * invP; gbar[5] = b[1] * invP;
        st->gam_old[0] = st->gam_old[1] = 0.0; st->alpha_old[0] = st->alpha_old[1] = 0.0; st->d_old[0] = st->d_old[1] = 0.0;
        st->stop2 = 0.0; st->done = 0; st->iters = 0; st->info = 0; st->pad = 0;
    }
}

// x = 0, r = b, p = s = 0, r32 = float(b) in the padded layout [3][hp][wp]
__global__ void __launch_bounds__(256) k_init(int w, int h, int wp, int hp, const double *__restrict__ b, double *__restrict__ x,
                                               double *__restrict__ r, double *__restrict__ p, double *__restrict__ s, float *__restrict__ r32)
{
    const unsigned int P = (unsigned int)w * h, k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const int i = (int)(k / (unsigned int)w), j = (int)(k - (unsigned int)i * w);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double bk = b[c * P + k];
        x[c * P + k] = 0.0; r[c * P + k] = bk; p[c * P + k] = 0.0; s[c * P + k] = 0.0;
        r32[((size_t)c * hp + i) * wp + j] = (float)bk;
    }
}

// zhat = (diag(alpha mu, alpha mu, lambda mu) + Gbar)^-1 rhat, in place on the padded spectra [3][hp][wp]
__global__ void __launch_bounds__(256) k_spectral(int w, int h, int wp, int hp, double alpha, double lam, const double *__restrict__ lam_x,
                                                   const double *__restrict__ lam_y, const double *__restrict__ gbar, float *__restrict__ T,
                                                   const int *skip)
{
    if (skip && *skip) return;
    const unsigned int P = (unsigned int)w * h, k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const int a = (int)(k / (unsigned int)w), b = (int)(k - (unsigned int)a * w);
    const double mu = lam_y[a] + lam_x[b];
    // Gbar is only positive semi-definite (fx = fy = 0 on a 2 x 2 image; a constant frame): a relative 1e-6 shift keeps
    // M positive definite at mu = 0 and is far below the accuracy a preconditioner needs
    const double reg = 1e-6 * (gbar[0] + gbar[3] + gbar[5]) + 1e-300;
    const double m00 = alpha * mu + gbar[0] + reg, m01 = gbar[1], m02 = gbar[2], m11 = alpha * mu + gbar[3] + reg, m12 = gbar[4],
                 m22 = lam * mu + gbar[5] + reg;
    const size_t i0 = ((size_t)a) * wp + b, cs = (size_t)hp * wp;
    const double r0 = T[i0], r1 = T[cs + i0], r2 = T[2 * cs + i0];
    // adjugate of the symmetric 3 x 3 matrix
    const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
    const double det = m00 * c00 + m01 * c01 + m02 * c02;
    const double id = det != 0.0 ? 1.0 / det : 0.0;
    T[i0] = (float)((c00 * r0 + c01 * r1 + c02 * r2) * id);
    T[cs + i0] = (float)((c01 * r0 + c11 * r1 + c12 * r2) * id);
    T[2 * cs + i0] = (float)((c02 * r0 + c12 * r1 + c22 * r2) * id);
}

// ---- even / odd folded transforms (w and h even): the DCT-II matrix satisfies C[k][n-1-i] = (-1)^k C[k][i], so a transform of
// length n is two transforms of length n/2 on the sums and differences x[i] +- x[n-1-i] (dct_kernels.cu, k_fold): half the
// flops and half the L2 traffic of the GEMMs, which is what they are bound by here.  The spectrum lives in the order
// "even frequencies, then odd" along both axes, in a [3][2 hq][2 wq] layout (hq, wq = half sizes padded to multiples of 4).

// fp64 n x n DCT matrix -> fp32 q x q: out[j][i] = C[2j+b][i] (transpose: out[i][j]) for i, j < n/2, zero padded
__global__ void k_pad_folded(int n, int q, int b, int transpose, const double *__restrict__ in, float *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= q * q) return;
    const int r = k / q, c = k - r * q, j = transpose ? c : r, i = transpose ? r : c;
    out[k] = (j < n / 2 && i < n / 2) ? (float)in[(size_t)(2 * j + b) * n + i] : 0.f;
}

// fold the middle axis of in[outer][n][inner] (outer stride so_in, row stride = inner_ld): out[b][outer][q][inner_ld] with
// rows >= n/2 zeroed (b = 0 sums, 1 differences; out outer stride = q * inner_ld, b stride = bs)
__global__ void __launch_bounds__(256) k_fold32(int outer, int n, int q, int inner, int inner_ld, long long so_in, const float *__restrict__ in,
                                                 float *__restrict__ out, long long bs, const int *skip)
{
    if (skip && *skip) return;
    const long long total = (long long)outer * q * inner_ld;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k % inner_ld);
        const long long oj = k / inner_ld;
        const int j = (int)(oj % q), o = (int)(oj / q);
        float e = 0.f, d = 0.f;
        if (j < n / 2 && i < inner) {
            const float a = in[o * so_in + (long long)j * inner_ld + i], c = in[o * so_in + (long long)(n - 1 - j) * inner_ld + i];
            e = a + c; d = a - c;
        }
        out[k] = e; out[bs + k] = d;
    }
}
// inverse: out[outer][j][i] = E + O, out[outer][n-1-j][i] = E - O for j < n/2
__global__ void __launch_bounds__(256) k_unfold32(int outer, int n, int q, int inner, int inner_ld, long long so_out, const float *__restrict__ in,
                                                   long long bs, float *__restrict__ out, const int *skip)
{
    if (skip && *skip) return;
    const long long total = (long long)outer * (n / 2) * inner;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k % inner);
        const long long oj = k / inner;
        const int j = (int)(oj % (n / 2)), o = (int)(oj / (n / 2));
        const long long src = ((long long)o * q + j) * inner_ld + i;
        const float e = in[src], d = in[bs + src];
        out[o * so_out + (long long)j * inner_ld + i] = e + d;
        out[o * so_out + (long long)(n - 1 - j) * inner_ld + i] = e - d;
    }
}
// the x axis is the innermost one: fold / unfold with the partner n-1-i inside a row.  in: rows x ld_in; folded: [b][rows][q]
__global__ void __launch_bounds__(256) k_fold32_x(int rows, int n, int q, int ld_in, const float *__restrict__ in, float *__restrict__ out,
                                                   const int *skip)
{
    if (skip && *skip) return;
    const long long total = (long long)rows * q, bs = total;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k % q);
        const long long r = k / q;
        float e = 0.f, d = 0.f;
        if (i < n / 2) { const float a = in[r * ld_in + i], c = in[r * ld_in + (n - 1 - i)]; e = a + c; d = a - c; }
        out[k] = e; out[bs + k] = d;
    }
}
__global__ void __launch_bounds__(256) k_unfold32_x(int rows, int n, int q, int ld_out, const float *__restrict__ in, float *__restrict__ out,
                                                     const int *skip)
{
    if (skip && *skip) return;
    const long long total = (long long)rows * (n / 2), bs = (long long)rows * q;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k % (n / 2));
        const long long r = k / (n / 2);
        const float e = in[r * q + i], d = in[bs + r * q + i];
        out[r * ld_out + i] = e + d;
        out[r * ld_out + (n - 1 - i)] = e - d;
    }
}

// k_spectral on the folded layout T[3][2 hq][2 wq]: position (by hq + jy, bx wq + jx) holds frequency (2 jy + by, 2 jx + bx)
__global__ void __launch_bounds__(256) k_spectral_folded(int w, int h, int wq, int hq, double alpha, double lam, const double *__restrict__ lam_x,
                                                          const double *__restrict__ lam_y, const double *__restrict__ gbar, float *__restrict__ T,
                                                          const int *skip)
{
    if (skip && *skip) return;
    const unsigned int P = (unsigned int)w * h, k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    const int hx = w / 2, hy = h / 2;
    const int ra = (int)(k / (unsigned int)w), rb = (int)(k - (unsigned int)ra * w);       // enumerate valid positions
    const int by = ra / hy, jy = ra - by * hy, bx = rb / hx, jx = rb - bx * hx;
    const double mu = lam_y[2 * jy + by] + lam_x[2 * jx + bx];
    const double reg = 1e-6 * (gbar[0] + gbar[3] + gbar[5]) + 1e-300;
    const double m00 = alpha * mu + gbar[0] + reg, m01 = gbar[1], m02 = gbar[2], m11 = alpha * mu + gbar[3] + reg, m12 = gbar[4],
                 m22 = lam * mu + gbar[5] + reg;
    const size_t i0 = ((size_t)(by * hq + jy)) * (2 * wq) + (bx * wq + jx), cs = (size_t)(2 * hq) * (2 * wq);
    const double r0 = T[i0], r1 = T[cs + i0], r2 = T[2 * cs + i0];
    const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
    const double det = m00 * c00 + m01 * c01 + m02 * c02;
    const double id = det != 0.0 ? 1.0 / det : 0.0;
    T[i0] = (float)((c00 * r0 + c01 * r1 + c02 * r2) * id);
    T[cs + i0] = (float)((c01 * r0 + c11 * r1 + c12 * r2) * id);
    T[2 * cs + i0] = (float)((c02 * r0 + c12 * r1 + c22 * r2) * id);
}

// w = A u (u in fp32, padded layout), partial sums (r.u, w.u, r.r) per block
__global__ void __launch_bounds__(256) k_stencil_dots(int w, int h, int wp, int hp, double alpha, double lam, const double *__restrict__ fx,
                                                       const double *__restrict__ fy, const double *__restrict__ f2,
                                                       const float *__restrict__ u32, const double *__restrict__ r, double *__restrict__ wv,
                                                       double *__restrict__ partials, const int *skip)
{
    __shared__ double red[32 * 3];
    if (skip && *skip) return;
    const unsigned int P = (unsigned int)w * h;
    const size_t cs = (size_t)hp * wp;
    double acc[3] = {0.0, 0.0, 0.0};
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < P; k += stride) {
        const int i = (int)(k / (unsigned int)w), j = (int)(k - (unsigned int)i * w);
        const size_t q = (size_t)i * wp + j;
        double uc[3], nl[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float *up = u32 + c * cs + q;
            const double ucc = (double)up[0];
            double sN = 0.0;
            if (i > 0) sN += ucc - (double)up[-wp];
            if (j > 0) sN += ucc - (double)up[-1];
            if (j < w - 1) sN += ucc - (double)up[1];
            if (i < h - 1) sN += ucc - (double)up[wp];
            uc[c] = ucc; nl[c] = sN;
        }
        const double fxk = fx[k], fyk = fy[k], f2k = f2[k];
        const double gp = fxk * uc[0] + fyk * uc[1] - f2k * uc[2];
        const double w0 = alpha * nl[0] + fxk * gp, w1 = alpha * nl[1] + fyk * gp, w2 = lam * nl[2] - f2k * gp;
        wv[k] = w0; wv[P + k] = w1; wv[2u * P + k] = w2;
        const double r0 = r[k], r1 = r[P + k], r2 = r[2u * P + k];
        acc[0] += r0 * uc[0] + r1 * uc[1] + r2 * uc[2];
        acc[1] += w0 * uc[0] + w1 * uc[1] + w2 * uc[2];
        acc[2] += r0 * r0 + r1 * r1 + r2 * r2;
    }
    block_sum<3>(acc, red);
    if (threadIdx.x == 0) { partials[3 * blockIdx.x] = acc[0]; partials[3 * blockIdx.x + 1] = acc[1]; partials[3 * blockIdx.x + 2] = acc[2]; }
}

// scalars of iteration `it` from the partials (every block, same order => same bits), stop test, vector updates
__global__ void __launch_bounds__(256) k_update(int w, int h, int wp, int hp, const double *__restrict__ partials, int nblk, int it, int maxiter,
                                                 double rtol, const float *__restrict__ u32, const double *__restrict__ wv, double *__restrict__ x,
                                                 double *__restrict__ r, double *__restrict__ p, double *__restrict__ s, float *__restrict__ r32,
                                                 GnDctState *st)
{
    __shared__ double red[32 * 3];
    if (st->done) return;                                // uniform over the grid: raised only by the last block to finish (below)
    double acc[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) { acc[0] += partials[3 * i]; acc[1] += partials[3 * i + 1]; acc[2] += partials[3 * i + 2]; }
    block_sum<3>(acc, red);
    const double gam = acc[0], del = acc[1], rho = acc[2];
    const int par = it & 1;
    const double stop2 = it == 0 ? (rtol * rtol) * rho : st->stop2;
    const bool converged = rho <= stop2 || rho == 0.0;
    const bool last = it + 1 >= maxiter;
    double alpha = 0.0, beta = 0.0, dk = 0.0;
    if (!converged) {
        beta = it == 0 ? 0.0 : gam / st->gam_old[par];
        dk = del - (beta * beta) * st->d_old[par];       // = delta - beta gamma / alpha_old
        alpha = gam / dk;
        const unsigned int P = (unsigned int)w * h;
        const size_t cs = (size_t)hp * wp;
        const unsigned int stride = gridDim.x * blockDim.x;
        for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < P; k += stride) {
            const int i = (int)(k / (unsigned int)w), j = (int)(k - (unsigned int)i * w);
            const size_t q = (size_t)i * wp + j;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const unsigned int kc = c * P + k;
                const double pv = (double)u32[c * cs + q] + beta * p[kc];
                const double sv = wv[kc] + beta * s[kc];
                p[kc] = pv; s[kc] = sv;
                x[kc] = x[kc] + alpha * pv;
                const double rv = r[kc] - alpha * sv;
                r[kc] = rv;
                r32[c * cs + q] = (float)rv;
            }
        }
    }
    // state for iteration it+1 goes to the other parity slot, so blocks of this launch that start later still read
    // the values of iteration it; `done` is raised by the LAST block to finish (ticket), after every block has read it
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(&st->pad, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        st->pad = 0;
        st->gam_old[par ^ 1] = gam; st->alpha_old[par ^ 1] = alpha; st->d_old[par ^ 1] = dk;
        if (it == 0) st->stop2 = stop2;
        if (converged) { st->done = 1; st->iters = it; st->info = 0; }
        else if (last) { st->done = 1; st->iters = it + 1; st->info = maxiter; }
        __threadfence();
    }
}

__global__ void __launch_bounds__(256) k_copy_out(unsigned int n, const double *__restrict__ x, double *__restrict__ u, double *__restrict__ v,
                                                   double *__restrict__ m, unsigned int P)
{
    const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < P) { u[k] = x[k]; v[k] = x[P + k]; m[k] = x[2u * P + k]; }
}

// tile choice: FOTO_GN_TILE = 0 (64 x 64, default), 1 (128 x 128), 2 (128 x 64).  Measured at 388x584 (B200, us per GEMM,
// X / Y transform): 64x64 15.0 / 11.1, 128x128 22.7 / 16.4, 128x64 slower than 64x64 too: the legacy mma.sync TF32 path
// delivers ~450 FLOP/clk/SM here, so fewer, larger CTAs (50 on 148 SMs) lose more than the halved L2 traffic gains.
// Splitting the 32-deep k-tile of a 64 x 64 block over two / four groups of warps (256 / 512 threads, accumulators added through
// shared memory) does not help either: 84.6 -> 92.2 / 97.1 us per PCG iteration at 388x584, 94.5 -> 99.3 / 107.5 at 480x640
// (profiles/r2_gn_ksplit.log); 32 x 64 and 64 x 32 tiles (370 CTAs, 2.5 per SM): 84.8 / 89.0 at 388x584, 107.0 / 107.7 at 480x640.
int tile_mode(int, int)
{
    static int forced = -2;
    if (forced == -2) { const char *e = getenv("FOTO_GN_TILE"); forced = e ? atoi(e) : -1; }
    return forced >= 0 ? forced : 0;
}
void gemm(cudaStream_t st, int M, int N, int K, const float *A, int lda, long long sA, const float *B, int ldb, long long sB,
          float *C, int ldc, long long sC, int batch, const int *skip)
{
    switch (tile_mode(M, N)) {
    case 1: GemmWide::launch(st, M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, batch, skip); break;
    case 2: GemmTall::launch(st, M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, batch, skip); break;
    default: GemmSmall::launch(st, M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, batch, skip); break;
    }
}

}  // namespace

size_t gn_dct_state_bytes() { return sizeof(GnDctState); }

// fp32 zero-padded copies of the DCT matrices of DctTables (Cx, CxT: wp x wp; Cy, CyT: hp x hp) and, for even w and h, the
// folded matrices (used from 1 M pixels on).
int gn_dct_prepare_tables(cudaStream_t st, const DctTables &tb, int w, int h, GnDctTables &out)
{
    const int wp = (w + 3) & ~3, hp = (h + 3) & ~3;
    // measured (ms per solve, folded / dense): 388x584 9.0 / 8.5, 380x420 7.4 / 6.4, 480x640 10.6 / 10.1, 1080x1920 58 / 81: below ~1 M
    // pixels the four extra launches of an application cost more than the halved GEMMs save (they are launch bound there)
    const char *env = getenv("FOTO_GN_FOLD");           // 0 / 1: force (A/B)
    const bool fold = !(w & 1) && !(h & 1) && (env ? atoi(env) != 0 : (long long)w * h >= (1ll << 20));
    if (out.base && out.w == w && out.h == h && out.fold == fold) return FOTO_OK;
    if (out.base) { CUDA_TRY(cudaFree(out.base)); out = GnDctTables(); }
    const int wq = (w / 2 + 3) & ~3, hq = (h / 2 + 3) & ~3;
    const size_t nx = (size_t)wp * wp, ny = (size_t)hp * hp, qx = (size_t)wq * wq, qy = (size_t)hq * hq;
    CUDA_TRY(cudaMalloc((void **)&out.base, (2 * nx + 2 * ny + (fold ? 4 * qx + 4 * qy : 0)) * sizeof(float)));
    out.Cx = out.base; out.CxT = out.Cx + nx; out.Cy = out.CxT + nx; out.CyT = out.Cy + ny;
    k_pad_matrix<<<(unsigned int)((nx + 255) / 256), 256, 0, st>>>(w, wp, tb.Cx, out.Cx);
    k_pad_matrix<<<(unsigned int)((nx + 255) / 256), 256, 0, st>>>(w, wp, tb.CxT, out.CxT);
    k_pad_matrix<<<(unsigned int)((ny + 255) / 256), 256, 0, st>>>(h, hp, tb.Cy, out.Cy);
    k_pad_matrix<<<(unsigned int)((ny + 255) / 256), 256, 0, st>>>(h, hp, tb.CyT, out.CyT);
    if (fold) {
        out.Ex = out.CyT + ny; out.ExT = out.Ex + 2 * qx; out.Ey = out.ExT + 2 * qx; out.EyT = out.Ey + 2 * qy;
        for (int b = 0; b < 2; b++) {
            k_pad_folded<<<(unsigned int)((qx + 255) / 256), 256, 0, st>>>(w, wq, b, 0, tb.Cx, out.Ex + b * qx);
            k_pad_folded<<<(unsigned int)((qx + 255) / 256), 256, 0, st>>>(w, wq, b, 1, tb.Cx, out.ExT + b * qx);
            k_pad_folded<<<(unsigned int)((qy + 255) / 256), 256, 0, st>>>(h, hq, b, 0, tb.Cy, out.Ey + b * qy);
            k_pad_folded<<<(unsigned int)((qy + 255) / 256), 256, 0, st>>>(h, hq, b, 1, tb.Cy, out.EyT + b * qy);
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(GemmSmall::prepare()); CUDA_TRY(GemmWide::prepare()); CUDA_TRY(GemmTall::prepare());       // per device: tables are per context
    out.w = w; out.h = h; out.wp = wp; out.hp = hp; out.fold = fold; out.wq = wq; out.hq = hq;
    return FOTO_OK;
}

// Enqueues the set-up of one solve; the iterations are enqueued by gn_dct_enqueue_iterations.
int gn_dct_begin(cudaStream_t st, const GnDctArgs &a)
{
    const unsigned int P = (unsigned int)a.w * a.h;
    const int blocks = (int)((P + 255) / 256) < 1184 ? (int)((P + 255) / 256) : 1184;
    k_gbar_partial<<<blocks, 256, 0, st>>>(P, a.fx, a.fy, a.f2, a.partials6);
    k_gbar_final<<<1, 256, 0, st>>>(a.partials6, blocks, 1.0 / (double)P, a.gbar, (GnDctState *)a.state);
    for (float *buf : {a.r32, a.t1, a.t2, a.u32})        // padding rows / columns stay zero
        CUDA_TRY(cudaMemsetAsync(buf, 0, a.tb->volume_floats() * sizeof(float), st));
    k_init<<<(P + 255) / 256, 256, 0, st>>>(a.w, a.h, a.tb->wp, a.tb->hp, a.b, a.x, a.r, a.p, a.s, a.r32);
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

int gn_dct_enqueue_iterations(cudaStream_t st, const GnDctArgs &a, int it0, int count, int *launches)
{
    const GnDctTables &tb = *a.tb;
    const int w = a.w, h = a.h, wp = tb.wp, hp = tb.hp;
    const unsigned int P = (unsigned int)w * h;
    const long long cs = (long long)hp * wp;
    const int *skip = &((GnDctState *)a.state)->done;
    const int pblocks = (int)((P + 255) / 256), nblk = pblocks < 592 ? pblocks : 592;
    for (int it = it0; it < it0 + count; it++) {
        if (tb.fold) {
            // u = M^-1 r with folded transforms: fold x, X GEMMs (sums | differences), fold y, Y GEMMs, 3x3 solve per frequency on
            // the permuted spectrum [3][2 hq][2 wq], and back
            const int wq = tb.wq, hq = tb.hq, wpp = 2 * wq, R = 3 * hp, fb = 148 * 8;
            const long long qx = (long long)wq * wq, qy = (long long)hq * hq, fyb = (long long)3 * hq * wpp, cs2 = (long long)2 * hq * wpp;
            k_fold32_x<<<fb, 256, 0, st>>>(R, w, wq, wp, a.r32, a.t1, skip);                                                     // t1 = FX[b][R][wq]
            gemm(st, R, wq, wq, a.t1, wq, (long long)R * wq, tb.ExT, wq, qx, a.t2, wpp, wq, 2, skip);                           // t2 = T1[R][wpp]
            k_fold32<<<fb, 256, 0, st>>>(3, h, hq, wpp, wpp, (long long)hp * wpp, a.t2, a.t1, fyb, skip);                       // t1 = FY[b][3][hq][wpp]
            for (int b = 0; b < 2; b++)       // (a second batch level inside the kernel cost the dense path 9 %: 94 -> 72 registers)
                gemm(st, hq, wpp, hq, tb.Ey + b * qy, hq, 0, a.t1 + b * fyb, wpp, (long long)hq * wpp, a.t2 + b * (long long)hq * wpp, wpp, cs2, 3, skip);   // t2 = T2[3][2hq][wpp]
            k_spectral_folded<<<pblocks, 256, 0, st>>>(w, h, wq, hq, a.alpha, a.lam, a.lam_x, a.lam_y, a.gbar, a.t2, skip);
            for (int b = 0; b < 2; b++)
                gemm(st, hq, wpp, hq, tb.EyT + b * qy, hq, 0, a.t2 + b * (long long)hq * wpp, wpp, cs2, a.t1 + b * fyb, wpp, (long long)hq * wpp, 3, skip); // t1 = GY[b][3][hq][wpp]
            k_unfold32<<<fb, 256, 0, st>>>(3, h, hq, wpp, wpp, (long long)hp * wpp, a.t1, fyb, a.t2, skip);                      // t2 = T1'[3][hp][wpp]
            gemm(st, R, wq, wq, a.t2, wpp, wq, tb.Ex, wq, qx, a.t1, wq, (long long)R * wq, 2, skip);                            // t1 = GX[b][R][wq]
            k_unfold32_x<<<fb, 256, 0, st>>>(R, w, wq, wp, a.t1, a.u32, skip);
            *launches += 6;
        } else {
        // u = M^-1 r:  T1 = R32 * CxT (rows of all three components at once), T2_c = Cy * T1_c, 3x3 solve per frequency,
        // T1_c = CyT * T2_c, U = T1 * Cx
        gemm(st, 3 * hp, wp, wp, a.r32, wp, 0, tb.CxT, wp, 0, a.t1, wp, 0, 1, skip);
        gemm(st, hp, wp, hp, tb.Cy, hp, 0, a.t1, wp, cs, a.t2, wp, cs, 3, skip);
        k_spectral<<<pblocks, 256, 0, st>>>(w, h, wp, hp, a.alpha, a.lam, a.lam_x, a.lam_y, a.gbar, a.t2, skip);
        gemm(st, hp, wp, hp, tb.CyT, hp, 0, a.t2, wp, cs, a.t1, wp, cs, 3, skip);
        gemm(st, 3 * hp, wp, wp, a.t1, wp, 0, tb.Cx, wp, 0, a.u32, wp, 0, 1, skip);
        }
        k_stencil_dots<<<nblk, 256, 0, st>>>(w, h, wp, hp, a.alpha, a.lam, a.fx, a.fy, a.f2, a.u32, a.r, a.wv, a.partials3, skip);
        k_update<<<nblk, 256, 0, st>>>(w, h, wp, hp, a.partials3, nblk, it, a.maxiter, a.rtol, a.u32, a.wv, a.x, a.r, a.p, a.s, a.r32,
                                       (GnDctState *)a.state);
        *launches += 7;
    }
    CUDA_TRY(cudaGetLastError());
    return FOTO_OK;
}

void gn_dct_copy_out(cudaStream_t st, const GnDctArgs &a, double *u, double *v, double *m)
{
    const unsigned int P = (unsigned int)a.w * a.h;
    k_copy_out<<<(P + 255) / 256, 256, 0, st>>>(3 * P, a.x, u, v, m, P);
}

void gn_dct_read_state(const void *host_copy, int *done, int *iters, int *info)
{
    const GnDctState *s = (const GnDctState *)host_copy;
    *done = s->done; *iters = s->iters; *info = s->info;
}

}  // namespace foto
