// TMA probe (round 2): fp64 tiled tensor maps, 3-D with halo and 4-D over components; shows that the innermost start
// coordinate of a box must be a multiple of 16 bytes (odd fp64 coordinate -> "illegal instruction") while negative /
// out-of-range coordinates in outer dimensions zero-fill.  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, int nbytes, double *out, int nout)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nbytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(sm)), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(smem_u32(sm)), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    const double *s = (const double *)sm;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
    void *p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaFree(0);
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    printf("entry point: %s qr %d ptr %p\n", cudaGetErrorString(e), (int)qr, p);
    Enc enc = (Enc)p;
    const int Nx = 584, Ny = 388, Nt = 4; const size_t P = (size_t)Nx * Ny, N = P * Nt;
    double *d, *out; cudaMalloc(&d, 3 * N * 8); cudaMalloc(&out, 4096 * 8);
    double *h = (double *)malloc(3 * N * 8);
    for (size_t i = 0; i < 3 * N; i++) h[i] = (double)i;
    cudaMemcpy(d, h, 3 * N * 8, cudaMemcpyHostToDevice);
    double ho[4096];
    int order[] = {0, 3, 4, 2, 5, 6, 7, 1};
    for (int oi = 0; oi < 8; oi++) { int test = order[oi];
        CUtensorMap tm; CUresult rc;
        int rank = 3, nb = 0, c[4] = {0, 0, 0, 0}, bx = 0, by = 0;
        if (test == 0 || test == 1 || test == 2 || test >= 5) {
            cuuint64_t dims[3] = {(cuuint64_t)Nx, (cuuint64_t)Ny, (cuuint64_t)Nt}; cuuint64_t st[2] = {(cuuint64_t)Nx * 8, P * 8};
            bx = test == 2 ? 64 : 66; by = 10;
            cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, es[3] = {1, 1, 1};
            rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            nb = bx * by * 8; c[0] = test == 1 ? -1 : 64; c[1] = test == 1 ? -1 : 8; c[2] = 1;
            if (test == 5) { c[0] = 64; c[1] = -1; }      // negative outer coordinate only
            if (test == 6) { c[0] = 63; c[1] = 8; }       // odd inner coordinate (8 mod 16 bytes)
            if (test == 7) { c[0] = -2; c[1] = 8; }       // negative, 16-byte aligned inner coordinate
        } else {
            rank = 4;
            cuuint64_t dims[4] = {(cuuint64_t)Nx, (cuuint64_t)Ny, (cuuint64_t)Nt, 3}; cuuint64_t st[3] = {(cuuint64_t)Nx * 8, P * 8, N * 8};
            bx = 64; by = 8;
            cuuint32_t box[4] = {64, 8, 1, (cuuint32_t)(test == 3 ? 1 : 3)}, es[4] = {1, 1, 1, 1};
            rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            nb = 64 * 8 * 8 * (test == 3 ? 1 : 3); c[0] = 64; c[1] = 8; c[2] = 1; c[3] = 0;
        }
        printf("test %d: encode rc %d, bytes %d\n", test, (int)rc, nb);
        if (rc) continue;
        if (rank == 3) k<3><<<1, 128, 32768>>>(tm, c[0], c[1], c[2], c[3], nb, out, nb / 8);
        else k<4><<<1, 128, 32768>>>(tm, c[0], c[1], c[2], c[3], nb, out, nb / 8);
        cudaError_t e2 = cudaDeviceSynchronize();
        printf("   run: %s\n", cudaGetErrorString(e2));
        if (e2) { cudaDeviceReset(); break; }
        cudaMemcpy(ho, out, nb, cudaMemcpyDeviceToHost);
        printf("   first words %.0f %.0f %.0f ... row1 %.0f (expect base %.0f)\n", ho[0], ho[1], ho[2], ho[bx], (double)((size_t)c[2] * P + (size_t)(c[1] < 0 ? 0 : c[1]) * Nx + (c[0] < 0 ? 0 : c[0])));
        if (rank == 4 && test == 4) printf("   comp1 first %.0f (expect %.0f)\n", ho[64 * 8], (double)(N + P + 8 * Nx + 64));
    }
    return 0;
}
