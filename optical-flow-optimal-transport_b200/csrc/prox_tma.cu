// prox_tma.cu -- K3 with TMA-staged inputs: grad_st phi + stepB + stepC + clamp + criterion partial sums
// (benamou_brenier.py:213-251), the same arithmetic as k_prox_dual (foto_kernels.cu), for whole volumes with even Nx.
//
// k_prox_dual is latency bound (every thread holds its next plane's four words in registers, 50 % occupancy, 69 % of
// the HBM peak at 1080x1920x16).  Here a CTA owns a TY x TX tile of the (y, x) plane and marches through t; one
// elected thread asks the TMA unit for
//   * the phi tile of plane n+1 with its halo (3-D tensor map, box (TX+4) x (TY+2) x 1 starting at (x0-2, y0-1):
//     the innermost start coordinate of a TMA box must be a multiple of 16 bytes -- an odd start raises "illegal
//     instruction", tools/tma_probe.cu -- so the x halo is two cells wide; the part of the box outside the image
//     is zero-filled by the hardware and never used), and
//   * the three mu components of plane n (4-D tensor map over [component][t][y][x], box TX x TY x 1 x 3)
// two steps ahead of their use; completion is signalled on an mbarrier per stage (expect_tx byte count), so loads
// cost neither registers nor issue slots and phi crosses HBM once.  phi planes n-1, n, n+1 live in a ring of four
// shared-memory buffers, mu in two stages; the six result words per cell are stored straight from registers
// (consecutive lanes = consecutive doubles).  46 KB of shared memory per CTA -> four CTAs of 256 threads per SM.
// Steps of consecutive tiles of a persistent CTA form one sequence, so the pipeline never drains between tiles.
#include <cuda.h>

#include "foto_kernels.cuh"
#include "prox_math.cuh"

namespace foto {

namespace {

constexpr int TX = 64, TY = 8, kThreads = 256, kRows = TY * TX / kThreads;   // rows per thread (2)
constexpr int kStages = 2, kRing = kStages + 2;
constexpr int kPhiPitch = TX + 4, kHaloX = 2;          // x halo of 2: TMA boxes start on 16-byte boundaries
constexpr int kPhiBox = kPhiPitch * (TY + 2) * 8;                           // 5440 bytes per TMA box
constexpr int kPhiBytes = (kPhiBox + 127) & ~127;                           // 5504
constexpr int kMuBytes = 3 * TY * TX * 8;                                     // 12288
constexpr int kSmemBytes = kRing * kPhiBytes + kStages * kMuBytes + 128;      // + alignment slack

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kThreads, 4) k_prox_dual_tma(const __grid_constant__ CUtensorMap tm_phi,
                                                              const __grid_constant__ CUtensorMap tm_mu, Dims d,
                                                              double *__restrict__ mu, double *__restrict__ q, double r,
                                                              double inv_r, int tiles_x, int ntiles,
                                                              double *__restrict__ partials)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ double red[64];
    __shared__ __align__(8) unsigned long long full[kStages];
    unsigned char *base = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    unsigned char *phi_ring = base, *mu_st = base + kRing * kPhiBytes;
    if (d.skip && *d.skip) return;
    const int tid = threadIdx.x, lx = tid % TX, ly0 = (tid / TX) * kRows;
    const int Nt = d.Nt;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nsteps = my_tiles * Nt;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // load group of step s: mu of (tile, n) and phi plane n+1 (plus plane 0 when a tile starts); thread 0 only
    auto issue = [&](int s) {
        const int ti = s / Nt, n = s - ti * Nt;
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        const int by = tile / tiles_x, bx = tile - by * tiles_x;
        const int x0 = bx * TX, y0 = by * TY;
        unsigned long long *bar = &full[s % kStages];
        const unsigned int bytes = (unsigned int)kMuBytes + (n + 1 < Nt ? (unsigned int)kPhiBox : 0u) + (n == 0 ? (unsigned int)kPhiBox : 0u);
        mbar_expect_tx(bar, bytes);
        tma_load_4d(mu_st + (s % kStages) * kMuBytes, &tm_mu, x0, y0, n, 0, bar);
        if (n == 0) tma_load_3d(phi_ring + (s % kRing) * kPhiBytes, &tm_phi, x0 - kHaloX, y0 - 1, 0, bar);
        if (n + 1 < Nt) tma_load_3d(phi_ring + ((s + 1) % kRing) * kPhiBytes, &tm_phi, x0 - kHaloX, y0 - 1, n + 1, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < kStages && s < nsteps; s++) issue(s);
    }

    double acc[2] = {0.0, 0.0};
    for (int s = 0; s < nsteps; s++) {
        const int ti = s / Nt, n = s - ti * Nt;
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        const int by = tile / tiles_x, bx = tile - by * tiles_x;
        const int x = bx * TX + lx, yb = by * TY + ly0;
        mbar_wait(&full[s % kStages], (unsigned int)((s / kStages) & 1));
        const double *pc = (const double *)(phi_ring + (s % kRing) * kPhiBytes);
        const double *pm = (const double *)(phi_ring + ((s + kRing - 1) % kRing) * kPhiBytes);
        const double *pp = (const double *)(phi_ring + ((s + 1) % kRing) * kPhiBytes);
        const double *ms = (const double *)(mu_st + (s % kStages) * kMuBytes);
        if (x < d.Nx) {
            // column x of the tile: rows ly0-1 .. ly0+kRows (tile coordinates are shifted by the halo)
            double col[kRows + 2];
#pragma unroll
            for (int j = 0; j < kRows + 2; j++) col[j] = pc[(ly0 + j) * kPhiPitch + lx + kHaloX];
#pragma unroll
            for (int j = 0; j < kRows; j++) {
                const int y = yb + j;
                if (y < d.Ny) {
                    const int ci = (ly0 + j + 1) * kPhiPitch + lx + kHaloX;
                    const double p_c = col[j + 1];
                    double gt;
                    if (n == 0) gt = pp[ci] - p_c;
                    else if (n == Nt - 1) gt = p_c - pm[ci];
                    else gt = 0.5 * pp[ci] - 0.5 * pm[ci];
                    double gx, gy;
                    if (x == 0) gx = pc[ci + 1] - p_c;
                    else if (x == d.Nx - 1) gx = p_c - pc[ci - 1];
                    else gx = 0.5 * pc[ci + 1] - 0.5 * pc[ci - 1];
                    if (y == 0) gy = col[j + 2] - p_c;
                    else if (y == d.Ny - 1) gy = p_c - col[j];
                    else gy = 0.5 * col[j + 2] - 0.5 * col[j];
                    const int mi = (ly0 + j) * TX + lx;
                    const double m0 = ms[mi], m1 = ms[TY * TX + mi], m2 = ms[2 * TY * TX + mi];
                    double qa, qb1, qb2;
                    project_K(gt + inv_r * m0, gx + inv_r * m1, gy + inv_r * m2, qa, qb1, qb2);
                    const unsigned int k = (unsigned int)n * d.P + (unsigned int)y * (unsigned int)d.Nx + (unsigned int)x;
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
        }
        __syncthreads();                                 // every thread is done with stage s and phi plane n-1
        if (tid == 0 && s + kStages < nsteps) issue(s + kStages);
    }
    block_sum<2>(acc, red);
    if (tid == 0) { partials[2 * blockIdx.x] = acc[0]; partials[2 * blockIdx.x + 1] = acc[1]; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

bool prox_tma_eligible(const Dims &d, const double *phi, const double *mu, const double *q)
{
    bool force = false;
    if (const char *e = getenv("FOTO_K3")) { if (!strcmp(e, "legacy")) return false; force = !strcmp(e, "tma"); }
    // below ~2 M cells the volume is L2 resident and a persistent CTA gets too few steps to fill its pipeline
    // (388x584x4: 27.6 us against 25.5 us for the register-marching kernel; 1080x1920x16: 0.513 against 0.585 ms)
    if (!force && (unsigned long long)d.N < (2ull << 20)) return false;
    if (d.n0 != 0 || d.gNt != d.Nt || d.cs != d.N) return false;          // time-slab views keep the register-marching kernel
    if ((d.Nx & 1) || d.Nt < 2) return false;                              // TMA strides must be multiples of 16 bytes
    if (((uintptr_t)phi | (uintptr_t)mu | (uintptr_t)q) & 15) return false;
    return encode_fn() != nullptr;
}

int launch_prox_dual_tma(cudaStream_t st, Dims d, const double *phi, double *mu, double *q, double r, double *partials,
                         int max_blocks, int num_sms, int *blocks_out)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return FOTO_ERR_CUDA; }
    CUtensorMap tm_phi, tm_mu;
    {
        cuuint64_t dims[3] = {(cuuint64_t)d.Nx, (cuuint64_t)d.Ny, (cuuint64_t)d.Nt};
        cuuint64_t strides[2] = {(cuuint64_t)d.Nx * 8, (cuuint64_t)d.P * 8};
        cuuint32_t box[3] = {kPhiPitch, TY + 2, 1}, es[3] = {1, 1, 1};
        CUresult rc = enc(&tm_phi, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)phi, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(phi) failed: %d", (int)rc); return FOTO_ERR_CUDA; }
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)d.Nx, (cuuint64_t)d.Ny, (cuuint64_t)d.Nt, 3};
        cuuint64_t strides[3] = {(cuuint64_t)d.Nx * 8, (cuuint64_t)d.P * 8, (cuuint64_t)d.cs * 8};
        cuuint32_t box[4] = {TX, TY, 1, 3}, es[4] = {1, 1, 1, 1};
        CUresult rc = enc(&tm_mu, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)mu, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(mu) failed: %d", (int)rc); return FOTO_ERR_CUDA; }
    }
    static bool attr_set[64] = {};                       // per device
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(k_prox_dual_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set[dev] = true;
    }
    const int tiles_x = (d.Nx + TX - 1) / TX, tiles_y = (d.Ny + TY - 1) / TY, ntiles = tiles_x * tiles_y;
    int blocks = 4 * num_sms;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks > ntiles) blocks = ntiles;
    k_prox_dual_tma<<<blocks, kThreads, kSmemBytes, st>>>(tm_phi, tm_mu, d, mu, q, r, 1.0 / r, tiles_x, ntiles, partials);
    *blocks_out = blocks;
    return FOTO_OK;
}

}  // namespace foto
