// Round-2 study of the grid all-reduce of the single-reduction CG kernel (cg_fused.cu): two values per CTA,
// one 448-thread CTA per SM with ~200 KB of dynamic shared memory (the real kernel's footprint).
// Every communicated word is self-validating (generation parity in the mantissa LSB), so no variant needs a fence,
// a sentinel reset or any ordering between different addresses.
//   A  two-hop root gather (round-1 shape), tagged words, single buffer
//   B  one-hop pull all-gather: every CTA publishes one 16-byte slot, warp 0 of every CTA polls all slots
//   C  thread-block clusters of CS CTAs: DSMEM pre-reduction to the cluster leader, leaders publish ncl slots and
//      poll all of them (one L2 hop), DSMEM broadcast back
//   D  clusters + root gather of the ncl cluster partials (two L2 hops, fewer words)
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench_allreduce2.cu -o /tmp/ub2
#include <cstdio>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long ldr(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
#ifdef NO_V2
__device__ __forceinline__ void str(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void ldr2(const unsigned long long *p, unsigned long long &a, unsigned long long &b) { a = ldr(p); b = ldr(p + 1); }
__device__ __forceinline__ void str2(unsigned long long *p, unsigned long long a, unsigned long long b) { str(p, a); str(p + 1, b); }
#else
__device__ __forceinline__ void ldr2(const unsigned long long *p, unsigned long long &a, unsigned long long &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ void str2(unsigned long long *p, unsigned long long a, unsigned long long b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }
#endif
// DSMEM: explicit shared::cluster state space (generic-address strong accesses to the shared window are slow)
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned int mapa(unsigned int a, int rank) { unsigned int r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ unsigned long long ldc(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ void stc(unsigned int a, unsigned long long v) { asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long tag(double v, unsigned int par) { return ((unsigned long long)__double_as_longlong(v) & ~1ull) | (par & 1u); }
__device__ __forceinline__ double val(unsigned long long b) { return __longlong_as_double((long long)b); }

__global__ void k_fill(unsigned long long *p, int n, unsigned long long v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

constexpr int kMaxB = 160;
// slots: [2 buffers][kMaxB][2 values] partials, then totals at 2*kMaxB*2

// ---- A: two-hop root gather, tagged, single buffer
__global__ void __launch_bounds__(448, 1) k_root(unsigned long long *slots, int iters, long long *cycles, double *out)
{
    extern __shared__ double dyn[];
    __shared__ double sh[2];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x, lane = tid & 31;
    double v0 = cta + 1.0, v1 = 2.0 * cta, t0v = 0, t1v = 0;
    unsigned long long *tot = slots + 4 * kMaxB;
    dyn[tid] = 0;
    long long t0 = clock64();
    for (unsigned int gen = 0; gen < (unsigned int)iters; gen++) {
        __syncthreads();
        const unsigned int par = gen & 1u;
        if (tid == 0) str2(slots + 2 * cta, tag(v0, par), tag(v1, par));
        if (cta == 0 && tid < 32) {
            unsigned long long a[5], b[5]; bool ready;
            do {
                ready = true;
#pragma unroll
                for (int k = 0; k < 5; k++) { int c = k * 32 + lane; ldr2(slots + 2 * (c < ncta ? c : lane), a[k], b[k]); }
#pragma unroll
                for (int k = 0; k < 5; k++) { int c = k * 32 + lane; ready = ready & ((a[k] & 1) == par) & ((b[k] & 1) == par); if (c >= ncta) { a[k] = 0; b[k] = 0; } }
            } while (!ready);
            double s0 = 0, s1 = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) { s0 += val(a[k]); s1 += val(b[k]); }
            for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(~0u, s0, o); s1 += __shfl_xor_sync(~0u, s1, o); }
            if (lane == 0) str2(tot, tag(s0, par), tag(s1, par));
        }
        if (tid == 0) {
            unsigned long long a, b;
            do { ldr2(tot, a, b); } while ((a & 1) != par || (b & 1) != par);
            sh[0] = val(a); sh[1] = val(b);
        }
        __syncthreads();
        t0v = sh[0]; t1v = sh[1];
        v0 = t0v * 1e-3 + cta; v1 = t1v * 1e-3 + 1;
    }
    if (tid == 0 && cta == 0) { *cycles = clock64() - t0; out[0] = t0v; out[1] = t1v; }
}

// ---- B: one-hop pull all-gather (double buffered by generation parity, tag = bit 1 of the generation)
__global__ void __launch_bounds__(448, 1) k_allpoll(unsigned long long *slots, int iters, long long *cycles, double *out)
{
    extern __shared__ double dyn[];
    __shared__ double sh[2];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x, lane = tid & 31;
    double v0 = cta + 1.0, v1 = 2.0 * cta, t0v = 0, t1v = 0;
    dyn[tid] = 0;
    long long t0 = clock64();
    for (unsigned int gen = 0; gen < (unsigned int)iters; gen++) {
        __syncthreads();
        const unsigned int par = (gen >> 1) & 1u;
        unsigned long long *buf = slots + (gen & 1u) * 2 * kMaxB;
        if (tid == 0) str2(buf + 2 * cta, tag(v0, par), tag(v1, par));
        if (tid < 32) {
            unsigned long long a[5], b[5]; bool ready;
            do {
                ready = true;
#pragma unroll
                for (int k = 0; k < 5; k++) { int c = k * 32 + lane; ldr2(buf + 2 * (c < ncta ? c : lane), a[k], b[k]); }
#pragma unroll
                for (int k = 0; k < 5; k++) { int c = k * 32 + lane; ready = ready & ((a[k] & 1) == par) & ((b[k] & 1) == par); if (c >= ncta) { a[k] = 0; b[k] = 0; } }
            } while (!ready);
            double s0 = 0, s1 = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) { s0 += val(a[k]); s1 += val(b[k]); }
            for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(~0u, s0, o); s1 += __shfl_xor_sync(~0u, s1, o); }
            if (lane == 0) { sh[0] = s0; sh[1] = s1; }
        }
        __syncthreads();
        t0v = sh[0]; t1v = sh[1];
        v0 = t0v * 1e-3 + cta; v1 = t1v * 1e-3 + 1;
    }
    if (tid == 0 && cta == 0) { *cycles = clock64() - t0; out[0] = t0v; out[1] = t1v; }
}


// ---- E: one-hop all-reduce by integer atomics.  A partial is cut into two 47-bit fixed-point limbs relative to a
// power-of-two scale every CTA derives from the previous total; each limb is added as (limb << 8) + 1 to a 64-bit
// accumulator that is never reset (the reader subtracts the word it saw two generations ago; double buffered), so
// the low byte of the difference counts arrivals and the word validates itself.  Integer addition is associative:
// every CTA gets the same bits whatever the arrival order.  STRIDE = distance between the four accumulators (words).
__device__ __forceinline__ void red_add(unsigned long long *p, unsigned long long v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
template <int STRIDE>
__global__ void __launch_bounds__(448, 1) k_atomic(unsigned long long *slots, int iters, long long *cycles, double *out)
{
    extern __shared__ double dyn[];
    __shared__ double sh[2];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x, lane = tid & 31;
    double v0 = cta + 1.0, v1 = 2.0 * cta, t0v = 12000.0, t1v = 170.0;   // "previous totals" seed the scale
    unsigned long long prev[2] = {0ull, 0ull};            // lane w (< 4): accumulator w of buffer 0 / 1 as last seen
    dyn[tid] = 0;
    long long t0 = clock64();
    for (unsigned int gen = 0; gen < (unsigned int)iters; gen++) {
        __syncthreads();
        if (tid < 4) {
            const int w = lane, vi = w >> 1;              // words 0,1: limbs of value 0; 2,3: value 1
            const double v = vi ? v1 : v0, pt = vi ? t1v : t0v;
            int e; frexp(pt, &e);                         // |pt| < 2^e
            const double q = ldexp(v, 47 - (e + 20));     // |q| < 2^47 as long as |v| < 2^20 |pt|
            const double hi = floor(q);
            const long long limb = (w & 1) ? (long long)floor(ldexp(q - hi, 47)) : (long long)hi;
            unsigned long long *acc = slots + ((gen & 1u) * 4 + w) * STRIDE;
            red_add(acc, ((unsigned long long)limb << 8) + 1ull);
            unsigned long long now, d;
            bool ready;
            do {
                now = ldr(acc);
                d = now - prev[gen & 1u];
                ready = (d & 0xFFull) == (unsigned long long)ncta;
            } while (!__all_sync(0xFu, ready));
            prev[gen & 1u] = now;
            const double part = (double)((long long)(d - (unsigned long long)ncta) >> 8);     // signed 56-bit sum of limbs
            const double lo = __shfl_down_sync(0xFu, part, 1);
            if ((w & 1) == 0) sh[vi] = ldexp(part + ldexp(lo, -47), (e + 20) - 47);
        }
        __syncthreads();
        t0v = sh[0]; t1v = sh[1];
        v0 = t0v * 1e-3 + cta; v1 = t1v * 1e-3 + 1;
    }
    if (tid == 0 && cta == 0) { *cycles = clock64() - t0; out[0] = t0v; out[1] = t1v; }
}

// ---- C / D: clusters.  MODE 0: leaders all-poll (one hop); MODE 1: root gather of cluster partials (two hops)
template <int CS, int MODE>
__global__ void __launch_bounds__(448, 1) k_cluster(unsigned long long *slots, int iters, long long *cycles, double *out)
{
    extern __shared__ double dyn[];
    __shared__ double sh[2];
    __shared__ __align__(16) unsigned long long inbox[2 * 8];      // leader: partials of the cluster's CTAs (tagged)
    __shared__ __align__(16) unsigned long long totbox[2];         // every CTA: totals (tagged), written by the leader
    cg::cluster_group cl = cg::this_cluster();
    const int tid = threadIdx.x, cta = blockIdx.x, lane = tid & 31;
    const int rank = (int)cl.block_rank(), ncl = gridDim.x / CS, cid = cta / CS;
    double v0 = cta + 1.0, v1 = 2.0 * cta, t0v = 0, t1v = 0;
    dyn[tid] = 0;
    if (tid < 16) inbox[tid] = ~0ull;
    if (tid < 2) totbox[tid] = ~0ull;
    const unsigned int lead_inbox = mapa(smem_u32(inbox), 0);
    unsigned long long *tot = slots + 4 * kMaxB;
    cl.sync();
    long long t0 = clock64();
    for (unsigned int gen = 0; gen < (unsigned int)iters; gen++) {
        __syncthreads();
        const unsigned int par1 = gen & 1u;                      // single-buffer words (DSMEM, root totals)
        const unsigned int par = MODE == 0 ? (gen >> 1) & 1u : par1;
        unsigned long long *buf = slots + (MODE == 0 ? (gen & 1u) : 0u) * 2 * kMaxB;
        if (rank != 0) {
            if (tid < 2) stc(lead_inbox + 8 * (2 * rank + tid), tag(tid ? v1 : v0, par1));
        } else if (tid < 32) {
            // leader: own partial + the cluster's, fixed order
            double c0 = v0, c1 = v1;
            if (lane == 0) {
#pragma unroll
                for (int r = 1; r < CS; r++) {
                    unsigned long long a, b;
                    do { a = ldc(inbox + 2 * r); b = ldc(inbox + 2 * r + 1); } while ((a & 1) != par1 || (b & 1) != par1);
                    c0 += val(a); c1 += val(b);
                }
                str2(buf + 2 * cid, tag(c0, par), tag(c1, par));
            }
            double s0 = 0, s1 = 0; bool have = false;
            if (MODE == 0 || cta == 0) {
                unsigned long long a[3], b[3]; bool ready;
                do {
                    ready = true;
#pragma unroll
                    for (int k = 0; k < 3; k++) { int c = k * 32 + lane; ldr2(buf + 2 * (c < ncl ? c : lane), a[k], b[k]); }
#pragma unroll
                    for (int k = 0; k < 3; k++) { int c = k * 32 + lane; ready = ready & ((a[k] & 1) == par) & ((b[k] & 1) == par); if (c >= ncl) { a[k] = 0; b[k] = 0; } }
                } while (!ready);
#pragma unroll
                for (int k = 0; k < 3; k++) { s0 += val(a[k]); s1 += val(b[k]); }
                for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(~0u, s0, o); s1 += __shfl_xor_sync(~0u, s1, o); }
                have = true;
                if (MODE == 1 && lane == 0) str2(tot, tag(s0, par1), tag(s1, par1));
            }
            if (MODE == 1 && !have) {
                unsigned long long a = 0, b = 0;
                if (lane == 0) { do { ldr2(tot, a, b); } while ((a & 1) != par1 || (b & 1) != par1); }
                s0 = val(__shfl_sync(~0u, a, 0)); s1 = val(__shfl_sync(~0u, b, 0));
            }
            // broadcast to the cluster through DSMEM (lane r writes to rank r)
            if (lane >= 1 && lane < CS) {
                const unsigned int rt = mapa(smem_u32(totbox), lane);
                stc(rt, tag(s0, par1)); stc(rt + 8, tag(s1, par1));
            }
            if (lane == 0) { sh[0] = val(tag(s0, par1)); sh[1] = val(tag(s1, par1)); }
        }
        if (rank != 0 && tid == 0) {
            unsigned long long a, b;
            do { a = ldc(totbox); b = ldc(totbox + 1); } while ((a & 1) != par1 || (b & 1) != par1);
            sh[0] = val(a); sh[1] = val(b);
        }
        __syncthreads();
        t0v = sh[0]; t1v = sh[1];
        v0 = t0v * 1e-3 + cta; v1 = t1v * 1e-3 + 1;
    }
    if (tid == 0 && cta == 0) { *cycles = clock64() - t0; out[0] = t0v; out[1] = t1v; }
    cl.sync();
}

template <typename K>
static void run_cluster(const char *name, K kern, int cs, int ncta, size_t smem, unsigned long long *slots, int iters, long long *cyc, double *res, bool coop)
{
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cs > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(448); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
    int maxcl = -1;
    cudaError_t eo = cudaOccupancyMaxActiveClusters(&maxcl, kern, &cfg);
    k_fill<<<8, 256>>>(slots, 2048, ~0ull);
    cudaError_t e = cudaSuccess;
    if (maxcl * cs >= ncta) e = cudaLaunchKernelEx(&cfg, kern, slots, iters, cyc, res);
    else { printf("%-44s cluster %d, %3d CTAs, coop %d: skipped, max active clusters %d (%s)\n", name, cs, ncta, (int)coop, maxcl, cudaGetErrorString(eo)); cudaGetLastError(); return; }
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h = 0; double r[2] = {0, 0}; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(r, res, 16, cudaMemcpyDeviceToHost);
    printf("%-44s cluster %d, %3d CTAs, coop %d: %6.0f cycles  [max clusters %d, %s/%s, totals %.3f %.3f]\n", name, cs, ncta, (int)coop, (double)h / iters, maxcl,
           cudaGetErrorString(e), cudaGetErrorString(e2), r[0], r[1]);
    cudaGetLastError();
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    unsigned long long *slots; long long *cyc; double *res;
    cudaMalloc(&slots, 2048 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&res, 16);
    int iters = 4000;
    const size_t smem = 200 * 1024;
    {   // warm the clocks up (an idle part runs its fabric slowly for the first milliseconds)
        unsigned long long *big; cudaMalloc(&big, 1 << 28);
        for (int k = 0; k < 400; k++) k_fill<<<(1 << 25) / 256, 256>>>(big, 1 << 25, (unsigned long long)k);
        cudaDeviceSynchronize(); cudaFree(big);
    }
    for (int rep = 0; rep < 2; rep++) {
    printf("---- pass %d\n", rep);
    for (int ncta : {144, 148}) {
        if (ncta > sms) continue;
        {
            cudaFuncSetAttribute(k_root, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_fill<<<8, 256>>>(slots, 2048, ~0ull);
            void *args[] = {&slots, &iters, &cyc, &res};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)k_root, dim3(ncta), dim3(448), args, smem, 0);
            cudaDeviceSynchronize();
            long long h; double r[2]; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(r, res, 16, cudaMemcpyDeviceToHost);
            printf("%-44s %3d CTAs: %6.0f cycles  [%s, totals %.3f %.3f]\n", "A two-hop root gather, tagged", ncta, (double)h / iters, cudaGetErrorString(e), r[0], r[1]);
        }
        {
            cudaFuncSetAttribute(k_allpoll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_fill<<<8, 256>>>(slots, 2048, ~0ull);
            void *args[] = {&slots, &iters, &cyc, &res};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)k_allpoll, dim3(ncta), dim3(448), args, smem, 0);
            cudaDeviceSynchronize();
            long long h; double r[2]; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(r, res, 16, cudaMemcpyDeviceToHost);
            printf("%-44s %3d CTAs: %6.0f cycles  [%s, totals %.3f %.3f]\n", "B one-hop pull all-gather", ncta, (double)h / iters, cudaGetErrorString(e), r[0], r[1]);
        }
        for (int stride : {1, 16}) {
            auto kern = stride == 1 ? k_atomic<1> : k_atomic<16>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_fill<<<8, 256>>>(slots, 2048, 0ull);
            void *args[] = {&slots, &iters, &cyc, &res};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)kern, dim3(ncta), dim3(448), args, smem, 0);
            cudaDeviceSynchronize();
            long long h; double r[2]; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(r, res, 16, cudaMemcpyDeviceToHost);
            printf("%-36s stride %2d %3d CTAs: %6.0f cycles  [%s, totals %.3f %.3f]\n", "E fixed-point atomics (one hop)", stride, ncta, (double)h / iters, cudaGetErrorString(e), r[0], r[1]);
        }
    }
    for (int coop : {1}) {
        run_cluster("C clusters, leaders all-poll (one hop)", k_cluster<2, 0>, 2, 144, smem, slots, iters, cyc, res, coop);
        run_cluster("C clusters, leaders all-poll (one hop)", k_cluster<2, 0>, 2, 148, smem, slots, iters, cyc, res, coop);
        run_cluster("D clusters, root gather (two hops)", k_cluster<2, 1>, 2, 144, smem, slots, iters, cyc, res, coop);
        run_cluster("C clusters, leaders all-poll (one hop)", k_cluster<4, 0>, 4, 144, smem, slots, iters, cyc, res, coop);
        run_cluster("C clusters, leaders all-poll (one hop)", k_cluster<4, 0>, 4, 132, smem, slots, iters, cyc, res, coop);
        run_cluster("D clusters, root gather (two hops)", k_cluster<4, 1>, 4, 132, smem, slots, iters, cyc, res, coop);
        run_cluster("C clusters, leaders all-poll (one hop)", k_cluster<8, 0>, 8, 128, smem, slots, iters, cyc, res, coop);
        run_cluster("C clusters, leaders all-poll (one hop)", k_cluster<8, 0>, 8, 144, smem, slots, iters, cyc, res, coop);
    }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
