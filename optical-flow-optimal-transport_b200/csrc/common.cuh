// common.cuh -- shared host/device helpers of libfoto_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "foto_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfoto_b200 is written for sm_100a (B200) only"
#endif

namespace foto {

void set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            foto::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                            cudaGetErrorString(e_));                                           \
            return FOTO_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

#define FOTO_TRY(expr)                                                                         \
    do {                                                                                       \
        int rc_ = (expr);                                                                      \
        if (rc_ != FOTO_OK) return rc_;                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------
// Grid-wide barrier fused with a deterministic all-reduce, for persistent kernels launched
// with cudaLaunchCooperativeKernel (co-residency guaranteed).
//
//   * every block reduces its values (shuffle butterfly + shared memory, fixed order),
//   * thread 0 publishes the block partial, fences, and arrives on a monotonically increasing
//     counter; it then spins with ld.acquire until all blocks of this generation have arrived,
//   * warp 0 of EVERY block re-reads all partials from L2 and sums them in the same fixed
//     order, so all blocks obtain bit-identical totals and take identical branches.
//
// Partials are double-buffered on the generation parity: a fast block may publish generation
// g+1 while a slow one still reads generation g, but g+2 cannot be published before every
// block has arrived at g+1, i.e. has finished reading g.
// ---------------------------------------------------------------------------------------
constexpr int kMaxBlocks = 1024;        // upper bound on persistent grid size
constexpr int kMaxVals = 4;             // values per all-reduce

struct SyncState {
    unsigned int *counter;              // zeroed by the host before every launch
    double *partials;                   // [2][kMaxVals][kMaxBlocks]
    int *error;                         // set to 1 by the watchdog
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV values; result valid in every thread.  smem: >= 32*NV doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
    __syncthreads();                                   // protect smem reuse across calls
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) smem[i * 32 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double t = lane < nwarp ? smem[i * 32 + lane] : 0.0;
        v[i] = warp_sum(t);
    }
}

// All-reduce NV doubles over the whole (co-resident) grid and synchronise it.
// gen: per-thread generation counter, starts at 0, identical in all threads.
template <int NV>
__device__ __forceinline__ void grid_allreduce(const SyncState &s, unsigned int &gen, double (&v)[NV],
                                               double *smem)
{
    static_assert(NV <= kMaxVals, "too many values");
    block_sum<NV>(v, smem);
    const unsigned int nblk = gridDim.x;
    double *part = s.partials + (size_t)(gen & 1u) * kMaxVals * kMaxBlocks;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) __stcg(&part[i * kMaxBlocks + blockIdx.x], v[i]);
        __threadfence();                               // release: block's writes + partial
        atomicAdd(s.counter, 1u);
        const unsigned int target = (gen + 1u) * nblk;
        long long t0 = clock64();
        while (ld_acquire_u32(s.counter) < target) {
            if (clock64() - t0 > 8000000000ll) { *s.error = 1; break; }   // ~4 s watchdog
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double t = 0.0;
            for (unsigned int b = threadIdx.x; b < nblk; b += 32) t += __ldcg(&part[i * kMaxBlocks + b]);
            t = warp_sum(t);
            if (threadIdx.x == 0) smem[i] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = smem[i];
    gen++;
}
#endif  // __CUDACC__

}  // namespace foto
