// Micro-benchmarks that size the on-chip CG kernel: fp64 pipe rate, shared-memory bandwidth,
// L2 round trip and grid-barrier latency on the machine at hand.  nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int OP> __global__ void k_fp64(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b); }
        if (OP == 1) { x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a); x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a); }
        if (OP == 2) { x0 = __dmul_rn(x0, a); x1 = __dmul_rn(x1, a); x2 = __dmul_rn(x2, a); x3 = __dmul_rn(x3, a); x4 = __dmul_rn(x4, a); x5 = __dmul_rn(x5, a); x6 = __dmul_rn(x6, a); x7 = __dmul_rn(x7, a); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_smem(double *out, int iters)
{
    extern __shared__ double s[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i;
    __syncthreads();
    double acc = 0;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) acc += s[(idx + k * 1024) & 8191];
        idx = (idx + 33) & 8191;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_gridsync(int iters, long long *cycles)
{
    cg::grid_group g = cg::this_grid();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) g.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

__global__ void k_l2lat(const int *chain, int iters, long long *cycles, int *sink)
{
    int p = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) p = __ldcg(chain + p);
    *cycles = clock64() - t0; *sink = p;
}

// ---- grid all-reduce variants (what the on-chip CG kernel needs twice per iteration) -------------
__device__ __forceinline__ unsigned long long ldr(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void fence_ar() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_sc() { asm volatile("fence.sc.gpu;" ::: "memory"); }
#define SENT 0x7FF8DEADBEEF0001ull

// mode bit0: arrive fence, bit1: root fence, bit2: waiter fence, bit3: use fence.sc instead of acq_rel
// mode bit4: counter barrier (atomicAdd + poll) instead of root gather
__global__ void k_allreduce(unsigned long long *slots, unsigned int *counter, int iters, int mode, long long *cycles, double *out)
{
    __shared__ double sh[2];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
    double val = cta + 1.0, total = 0.0;
    long long t0 = clock64();
    for (unsigned int gen = 0; gen < (unsigned int)iters; gen++) {
        __syncthreads();
        if (mode & 16) {
            if (tid == 0) {
                str(slots + (gen & 1) * 1024 + cta, (unsigned long long)__double_as_longlong(val));
                if (mode & 1) { if (mode & 8) fence_sc(); else fence_ar(); }
                atomicAdd(counter, 1u);
                const unsigned int target = (gen + 1) * ncta;
                unsigned int c;
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(counter) : "memory"); } while (c < target);
                if (mode & 4) { if (mode & 8) fence_sc(); else fence_ar(); }
            }
            __syncthreads();
            if (tid < 32) {
                double s = 0.0;
                unsigned long long v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) { int b = k * 32 + tid; v[k] = b < ncta ? ldr(slots + (gen & 1) * 1024 + b) : 0ull; }
#pragma unroll
                for (int k = 0; k < 8; k++) s += __longlong_as_double((long long)v[k]);
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (tid == 0) sh[0] = s;
            }
            __syncthreads();
            total = sh[0];
        } else {
            if (tid == 0) {
                str(slots + ((gen + 1) % 3) * 1024 + cta, SENT);
                if (mode & 1) { if (mode & 8) fence_sc(); else fence_ar(); }
                str(slots + (gen % 3) * 1024 + cta, (unsigned long long)__double_as_longlong(val));
            }
            if (cta == 0 && tid < 32) {
                unsigned long long v[8]; bool ready;
                do {
                    ready = true;
#pragma unroll
                    for (int k = 0; k < 8; k++) { int b = k * 32 + tid; v[k] = b < ncta ? ldr(slots + (gen % 3) * 1024 + b) : 0ull; ready = ready && v[k] != SENT; }
                } while (!ready);
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) s += __longlong_as_double((long long)v[k]);
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (tid == 0) {
                    str(slots + 3072 + ((gen + 1) % 3) * 16, SENT);
                    if (mode & 2) { if (mode & 8) fence_sc(); else fence_ar(); }
                    str(slots + 3072 + (gen % 3) * 16, (unsigned long long)__double_as_longlong(s));
                }
            }
            if (tid == 0) {
                unsigned long long b;
                while ((b = ldr(slots + 3072 + (gen % 3) * 16)) == SENT) {}
                if (mode & 4) { if (mode & 8) fence_sc(); else fence_ar(); }
                sh[0] = __longlong_as_double((long long)b);
            }
            __syncthreads();
            total = sh[0];
        }
        val = total * 1e-3 + cta;
    }
    if (tid == 0 && cta == 0) { *cycles = clock64() - t0; *out = total; }
}

// pure streaming with R read streams and W write streams of n doubles each (what K3 does: R=4, W=6)
template <int R, int W> __global__ void __launch_bounds__(256) k_stream(const double *__restrict__ in, double *__restrict__ out, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < R; r++) s += in[r * n + k];
#pragma unroll
        for (int w = 0; w < W; w++) out[w * n + k] = s + w;
    }
}

__global__ void k_fill(unsigned long long *p, int n, unsigned long long v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("%s, %d SMs, %d MHz, L2 %d MB\n", prop.name, sms, prop.clockRate / 1000, prop.l2CacheSize >> 20);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const char *names[3] = {"DFMA", "DADD", "DMUL"};
    for (int op = 0; op < 3; op++) {
        int iters = 20000;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (op == 0) k_fp64<0><<<sms * 2, 1024>>>(out, iters, 1.0000001, 1e-9);
            if (op == 1) k_fp64<1><<<sms * 2, 1024>>>(out, iters, 1.0000001, 1e-9);
            if (op == 2) k_fp64<2><<<sms * 2, 1024>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double ops = (double)sms * 2 * 1024 * iters * 8;
        printf("%s: %.2f Tops/s  = %.1f lanes/clk/SM at %d MHz nominal\n", names[op], ops / ms / 1e9,
               ops / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1000);
    }
    {
        int iters = 4000;
        cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            k_smem<<<sms, 1024, 65536>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double bytes = (double)sms * 1024 * iters * 8 * 8;
        printf("LDS.64: %.1f TB/s = %.1f B/clk/SM\n", bytes / ms / 1e9, bytes / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
    }
    {
        long long *cyc; cudaMalloc(&cyc, 8);
        int iters = 2000;
        for (int bs : {128, 1024}) {
            void *args[] = {&iters, &cyc};
            cudaLaunchCooperativeKernel((void *)k_gridsync, dim3(sms), dim3(bs), args, 0, 0);
            cudaEventRecord(e0);
            cudaLaunchCooperativeKernel((void *)k_gridsync, dim3(sms), dim3(bs), args, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("cg grid.sync (%d CTAs x %d thr): %.2f us = %.0f cycles per sync\n", sms, bs, 1e3 * ms / iters, (double)h / iters);
        }
    }
    {
        int n = 1 << 20, *chain, *sink; long long *cyc;
        cudaMalloc(&chain, n * 4); cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
        int *h = new int[n];
        for (int i = 0; i < n; i++) h[i] = (i + 4099) % n;
        cudaMemcpy(chain, h, n * 4, cudaMemcpyHostToDevice);
        k_l2lat<<<1, 1>>>(chain, 20000, cyc, sink);
        k_l2lat<<<1, 1>>>(chain, 20000, cyc, sink);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("L2 dependent-load latency: %.0f cycles\n", (double)c / 20000);
    }
    {
        unsigned long long *slots; unsigned int *counter; long long *cyc; double *res;
        cudaMalloc(&slots, 4096 * 8); cudaMalloc(&counter, 4); cudaMalloc(&cyc, 8); cudaMalloc(&res, 8);
        int iters = 3000;
        const int modes[] = {7, 0, 1, 3, 5, 15, 16 + 5, 16, 16 + 13};
        for (int bs : {1024}) for (int mode : modes) {
            k_fill<<<16, 256>>>(slots, 4096, SENT); cudaMemset(counter, 0, 4);
            void *args[] = {&slots, &counter, &iters, (void *)&mode, &cyc, &res};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)k_allreduce, dim3(sms), dim3(bs), args, 0, 0);
            cudaDeviceSynchronize();
            long long h; double r; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&r, res, 8, cudaMemcpyDeviceToHost);
            printf("allreduce mode %2d (%s; fences: arrive %d root %d wait %d %s): %.0f cycles  [%s, total %.3f]\n", mode,
                   (mode & 16) ? "counter+partials" : "root gather", mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, (mode & 8) ? "sc" : "acq_rel",
                   (double)h / iters, cudaGetErrorString(e), r);
        }
    }
    {
        const size_t n = 33177600;     // 1080 x 1920 x 16
        double *in, *o; cudaMalloc(&in, n * 8 * 8); cudaMalloc(&o, n * 8 * 8);
        cudaMemset(in, 0, n * 8 * 8);
        auto run = [&](auto kern, int R, int W, const char *name) {
            for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); kern<<<148 * 16, 256>>>(in, o, n); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
            printf("stream %s: %d read + %d write streams of 265 MB: %.3f ms = %.0f GB/s\n", name, R, W, ms, (R + W) * n * 8 / ms / 1e6);
        };
        run(k_stream<1, 1>, 1, 1, "copy");
        run(k_stream<4, 6>, 4, 6, "K3-like");
        run(k_stream<6, 1>, 6, 1, "K1-like");
        run(k_stream<6, 4>, 6, 4, "CG-like");
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
