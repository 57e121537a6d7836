// Die-aware grid all-reduce on the two-die B200 (round 2).  tools/ubench_die.cu shows that a polled flag costs ~400 cycles
// more per hop when its L2 home is on the other die than the poller (store -> visible: 477 cycles one way with both SMs
// and the word on one die, 643 across the dies with the word homed on the poller's side, ~1 050 otherwise), and that the
// home changes with the address at 2 KB granularity.  Variant A (what grid_sync.cuh did) puts all words in one buffer
// wherever they fall.  Variant F: 16 candidate 2 KB granules; every CTA classifies them (own die / other die) by a self
// ping-pong; the root gathers the partials in two granules of its own die and stores the totals into one granule of each
// die; every waiter polls the copy on ITS die.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench_allreduce_die.cu -o /tmp/ubd
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ldr(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void ldr2(const unsigned long long *p, unsigned long long &a, unsigned long long &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ void str2(unsigned long long *p, unsigned long long a, unsigned long long b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }
__device__ __forceinline__ unsigned long long tag(double v, unsigned int par) { return ((unsigned long long)__double_as_longlong(v) & ~1ull) | (par & 1u); }
__device__ __forceinline__ double val(unsigned long long b) { return __longlong_as_double((long long)b); }
__global__ void k_fill(unsigned long long *p, int n, unsigned long long v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

constexpr int kGran = 16, kGranWords = 256, kSlotsPerGran = 120, kTotOff = 240, kChoiceOff = 244, kProbeOff = 246;

// ---- A: two-hop root gather, one buffer (partials at words [0, 2 ncta), totals at word 1024)
__global__ void __launch_bounds__(448, 1) k_root(unsigned long long *slots, int iters, long long *cycles, double *out)
{
    extern __shared__ double dyn[];
    __shared__ double sh[2];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x, lane = tid & 31;
    double v0 = cta + 1.0, v1 = 2.0 * cta, t0v = 0, t1v = 0;
    unsigned long long *tot = slots + 1024;
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

// which granules are homed on this SM's die?  Self ping-pong: store a new value to a private word of the granule and poll
// it back (the store has to reach the L2 home, the strong load has to come back from it): ~480 cycles with the home on
// this die, ~900 on the other one.  (The latency of a strong load alone does not tell: 280 against 310 cycles.)
__device__ unsigned int local_mask(unsigned long long *base, int cta, int *lat)
{
    int t[kGran], lo = 1 << 30, hi = 0;
    for (int g = 0; g < kGran; g++) {
        unsigned long long *p = base + g * kGranWords + cta;
        int m = 1 << 30;
        for (int rep = 0; rep < 6; rep++) {
            const long long c0 = clock64();
            str(p, 0x5000ull + rep);
            while (ldr(p) != 0x5000ull + rep) { }
            const int d = (int)(clock64() - c0);
            m = d < m ? d : m;
        }
        t[g] = m; if (lat) lat[g] = m;
        lo = m < lo ? m : lo; hi = m > hi ? m : hi;
    }
    unsigned int mask = 0;
    for (int g = 0; g < kGran; g++) if (2 * t[g] < lo + hi) mask |= 1u << g;
    return mask;
}

// ---- F: die-aware placement, self-calibrating
__global__ void __launch_bounds__(448, 1) k_root_die(unsigned long long *base, int iters, long long *cycles, double *out, int *lat_out)
{
    extern __shared__ double dyn[];
    __shared__ double sh[2];
    __shared__ int cfg[5];       // my totals granule, root's two partial granules, the two totals granules
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x, lane = tid & 31;
    double v0 = cta + 1.0, v1 = 2.0 * cta, t0v = 0, t1v = 0;
    dyn[tid] = 0;
    if (tid == 0) {
        const unsigned int mask = local_mask(base, cta, (cta == 0 || cta == ncta - 1) ? lat_out + (cta == 0 ? 0 : kGran) : nullptr);
        if (cta == 0) {
            // partials in the root's first two local granules, totals in its third local granule and in its first remote one
            int loc[3] = {0, 1, 2}, nl = 0, rem = -1;
            for (int g = 0; g < kGran; g++) { if ((mask >> g) & 1u) { if (nl < 3) loc[nl++] = g; } else if (rem < 0) rem = g; }
            if (rem < 0) rem = loc[2];
            const unsigned long long ch = 0xC0DE00000000ull | (unsigned long long)loc[0] | ((unsigned long long)loc[1] << 8) |
                                          ((unsigned long long)loc[2] << 16) | ((unsigned long long)rem << 24);
            for (int g = 0; g < kGran; g++) str(base + g * kGranWords + kChoiceOff, ch);
        }
        unsigned long long ch;
        do { ch = ldr(base + kChoiceOff); } while ((ch >> 32) != 0xC0DEull);
        const int ga = (int)(ch & 255), gb = (int)((ch >> 8) & 255), tA = (int)((ch >> 16) & 255), tB = (int)((ch >> 24) & 255);
        cfg[1] = ga; cfg[2] = gb; cfg[3] = tA; cfg[4] = tB;
        cfg[0] = ((mask >> tA) & 1u) ? tA : tB;              // the copy of the totals on this SM's die
    }
    __syncthreads();
    const int gw = cfg[0], ga = cfg[1], gb = cfg[2];
    unsigned long long *myslot = base + (cta < kSlotsPerGran ? ga : gb) * kGranWords + 2 * (cta % kSlotsPerGran);
    const unsigned long long *mytot = base + gw * kGranWords + kTotOff;
    long long t0 = clock64();
    for (unsigned int gen = 0; gen < (unsigned int)iters; gen++) {
        __syncthreads();
        const unsigned int par = gen & 1u;
        if (tid == 0) str2(myslot, tag(v0, par), tag(v1, par));
        if (cta == 0 && tid < 32) {
            unsigned long long a[5], b[5]; bool ready;
            do {
                ready = true;
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    int c = k * 32 + lane; if (c >= ncta) c = lane;
                    ldr2(base + (c < kSlotsPerGran ? ga : gb) * kGranWords + 2 * (c % kSlotsPerGran), a[k], b[k]);
                }
#pragma unroll
                for (int k = 0; k < 5; k++) { int c = k * 32 + lane; ready = ready & ((a[k] & 1) == par) & ((b[k] & 1) == par); if (c >= ncta) { a[k] = 0; b[k] = 0; } }
            } while (!ready);
            double s0 = 0, s1 = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) { s0 += val(a[k]); s1 += val(b[k]); }
            for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(~0u, s0, o); s1 += __shfl_xor_sync(~0u, s1, o); }
            if (lane < 2) str2(base + cfg[3 + lane] * kGranWords + kTotOff, tag(s0, par), tag(s1, par));
        }
        if (tid == 0) {
            unsigned long long a, b;
            do { ldr2(mytot, a, b); } while ((a & 1) != par || (b & 1) != par);
            sh[0] = val(a); sh[1] = val(b);
        }
        __syncthreads();
        t0v = sh[0]; t1v = sh[1];
        v0 = t0v * 1e-3 + cta; v1 = t1v * 1e-3 + 1;
    }
    if (tid == 0 && cta == 0) { *cycles = clock64() - t0; out[0] = t0v; out[1] = t1v; }
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    unsigned long long *raw, *slots; long long *cyc; double *res; int *lat;
    const int words = (kGran + 1) * kGranWords;
    cudaMalloc(&raw, (words + 512) * 8); cudaMalloc(&cyc, 8); cudaMalloc(&res, 16); cudaMalloc(&lat, 2 * kGran * sizeof(int));
    slots = (unsigned long long *)(((size_t)raw + 2047) & ~(size_t)2047);           // 2 KB aligned
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k_root, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_root_die, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int iters = 4000;
    for (int pass = 0; pass < 2; pass++)
        for (int ncta : {144, 148}) {
            for (int shift = 0; shift < 3; shift++) {             // variant A at three buffer positions (different homes)
                unsigned long long *s = slots + shift * 3 * kGranWords;
                k_fill<<<(words + 255) / 256, 256>>>(slots, words, ~0ull);
                void *args[] = {&s, &iters, &cyc, &res};
                cudaError_t e = cudaLaunchCooperativeKernel((void *)k_root, dim3(ncta), dim3(448), args, smem, 0);
                cudaDeviceSynchronize();
                long long h; double r[2]; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(r, res, 16, cudaMemcpyDeviceToHost);
                printf("A one buffer (at granule %d)          %d CTAs: %6.0f cycles  [%s, totals %.3f %.3f]\n", shift * 3, ncta, (double)h / iters, cudaGetErrorString(e), r[0], r[1]);
            }
            k_fill<<<(words + 255) / 256, 256>>>(slots, words, ~0ull);
            void *args[] = {&slots, &iters, &cyc, &res, &lat};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)k_root_die, dim3(ncta), dim3(448), args, smem, 0);
            cudaDeviceSynchronize();
            long long h; double r[2]; int hl[2 * kGran];
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(r, res, 16, cudaMemcpyDeviceToHost); cudaMemcpy(hl, lat, sizeof(hl), cudaMemcpyDeviceToHost);
            printf("F die-aware, self-calibrating         %d CTAs: %6.0f cycles  [%s, totals %.3f %.3f]\n", ncta, (double)h / iters, cudaGetErrorString(e), r[0], r[1]);
            printf("   self ping-pong cycles per granule, CTA 0:     "); for (int g = 0; g < kGran; g++) printf(" %4d", hl[g]); printf("\n");
            printf("   self ping-pong cycles per granule, last CTA:  "); for (int g = 0; g < kGran; g++) printf(" %4d", hl[kGran + g]); printf("\n");
        }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
