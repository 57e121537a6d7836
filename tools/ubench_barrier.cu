// Grid all-reduce latency study for the on-chip CG kernels (cg_onchip.cu, gn_onchip.cu).
// Variants of the root-gather barrier: number of polling warps in the root CTA, number of polling warps in the
// waiting CTAs (first one to see the result raises a shared-memory flag), number of replicas of the broadcast
// word, and a one-way signal ping-pong as the floor.   nvcc -O3 -arch=sm_100a tools/ubench_barrier.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ldr(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
#define SENT 0x7FF8DEADBEEF0001ull
constexpr int kBcast = 3072;      // slots: [3][1024] partials, then [3][REPMAX=32][16 words] broadcast lines

__global__ void k_fill(unsigned long long *p, int n, unsigned long long v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

// RW root warps, WW waiter warps, REP broadcast replicas
__global__ void __launch_bounds__(512) k_allreduce(unsigned long long *slots, int iters, int RW, int WW, int REP, long long *cycles, double *out)
{
    __shared__ double sh[2];
    __shared__ volatile int flag[4];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x, warp = tid >> 5, lane = tid & 31;
    double val = cta + 1.0, total = 0.0;
    if (tid < 4) flag[tid] = 0;
    __syncthreads();
    long long t0 = clock64();
    for (unsigned int gen = 0; gen < (unsigned int)iters; gen++) {
        __syncthreads();
        if (tid == 0) {
            str(slots + ((gen + 1) % 3) * 1024 + cta, SENT);
            str(slots + (gen % 3) * 1024 + cta, (unsigned long long)__double_as_longlong(val));
        }
        if (cta == 0 && warp < RW) {
            unsigned long long v[8]; bool ready;
            do {
                ready = true;
#pragma unroll
                for (int k = 0; k < 8; k++) { int b = k * 32 + lane; v[k] = b < ncta ? ldr(slots + (gen % 3) * 1024 + b) : 0ull; ready = ready && v[k] != SENT; }
            } while (!ready);
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) s += __longlong_as_double((long long)v[k]);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane < REP) {
                str(slots + kBcast + (((gen + 1) % 3) * 32 + lane) * 16, SENT);
                str(slots + kBcast + ((gen % 3) * 32 + lane) * 16, (unsigned long long)__double_as_longlong(s));
            }
        }
        if (warp < WW && lane == 0) {
            const unsigned long long *p = slots + kBcast + ((gen % 3) * 32 + cta % REP) * 16;
            unsigned long long b;
            const int want = (int)(gen & 1) + 1;           // flag value for this generation
            if (WW == 1) {
                while ((b = ldr(p)) == SENT) {}
                sh[0] = __longlong_as_double((long long)b);
            } else {
                // stagger the pollers, stop as soon as any of them has seen the value
                for (int d = 0; d < warp * 40; d++) asm volatile("");
                while (true) {
                    b = ldr(p);
                    if (b != SENT) { sh[0] = __longlong_as_double((long long)b); __threadfence_block(); flag[0] = want; break; }
                    if (flag[0] == want) break;
                }
            }
        }
        __syncthreads();
        total = sh[0];
        val = total * 1e-3 + cta;
    }
    if (tid == 0 && cta == 0) { *cycles = clock64() - t0; *out = total; }
}


// one-hop variant: every CTA pushes its partial into a private inbox of every other CTA ([3][ncta][ncta] words) and
// polls only its own inbox; all CTAs sum the same values in the same order.
__global__ void __launch_bounds__(512) k_allgather_push(unsigned long long *inbox, int iters, int fence, long long *cycles, double *out)
{
    __shared__ double sh[2];
    const int tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x, lane = tid & 31;
    double val = cta + 1.0, total = 0.0;
    long long t0 = clock64();
    for (unsigned int gen = 0; gen < (unsigned int)iters; gen++) {
        __syncthreads();
        if (tid < 32) {
            unsigned long long *nxt = inbox + (size_t)((gen + 1) % 3) * ncta * ncta, *cur = inbox + (size_t)(gen % 3) * ncta * ncta;
            for (int d = lane; d < ncta; d += 32) str(nxt + (size_t)d * ncta + cta, SENT);
            if (fence) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            const unsigned long long bits = (unsigned long long)__double_as_longlong(val);
            for (int d = lane; d < ncta; d += 32) str(cur + (size_t)d * ncta + cta, bits);
            const unsigned long long *mine = cur + (size_t)cta * ncta;
            unsigned long long v[8]; bool ready;
            do {
                ready = true;
#pragma unroll
                for (int k = 0; k < 8; k++) { int b = k * 32 + lane; v[k] = b < ncta ? ldr(mine + b) : 0ull; ready = ready && v[k] != SENT; }
            } while (!ready);
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) s += __longlong_as_double((long long)v[k]);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sh[0] = s;
        }
        __syncthreads();
        total = sh[0];
        val = total * 1e-3 + cta;
    }
    if (tid == 0 && cta == 0) { *cycles = clock64() - t0; *out = total; }
}

// one-way signal latency: CTA 0 and CTA `peer` ping-pong a word through L2
__global__ void k_pingpong(unsigned long long *w, int iters, int peer, long long *cycles)
{
    const int cta = blockIdx.x;
    if (threadIdx.x != 0 || (cta != 0 && cta != peer)) return;
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        if (cta == 0) { str(w, (unsigned long long)(2 * i - 1)); while (ldr(w + 16) != (unsigned long long)(2 * i)) {} }
        else { while (ldr(w) != (unsigned long long)(2 * i - 1)) {} str(w + 16, (unsigned long long)(2 * i)); }
    }
    if (cta == 0) *cycles = clock64() - t0;
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    unsigned long long *slots; long long *cyc; double *res;
    cudaMalloc(&slots, 8192 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&res, 8);
    int iters = 3000;
    {
        for (int peer : {1, 2, 37, 73, 74, 100, 147}) {
            cudaMemset(slots, 0, 8192 * 8);
            void *args[] = {&slots, &iters, &peer, &cyc};
            cudaLaunchCooperativeKernel((void *)k_pingpong, dim3(sms), dim3(32), args, 0, 0);
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("ping-pong CTA 0 <-> CTA %3d: %.0f cycles per round trip (2 one-way signals)\n", peer, (double)h / iters);
        }
    }
    for (int bs : {512}) for (int RW : {1, 4}) for (int WW : {1, 2}) for (int REP : {1, 4}) {
        k_fill<<<32, 256>>>(slots, 8192, SENT);
        void *args[] = {&slots, &iters, &RW, &WW, &REP, &cyc, &res};
        cudaError_t e = cudaLaunchCooperativeKernel((void *)k_allreduce, dim3(sms), dim3(bs), args, 0, 0);
        cudaDeviceSynchronize();
        long long h; double r; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&r, res, 8, cudaMemcpyDeviceToHost);
        printf("allreduce root warps %d, waiter warps %d, bcast replicas %2d: %.0f cycles  [%s, total %.3f]\n", RW, WW, REP, (double)h / iters, cudaGetErrorString(e), r);
    }
    {
        unsigned long long *inbox; const int n = 3 * sms * sms;
        cudaMalloc(&inbox, (size_t)n * 8);
        for (int fence : {0, 1}) {
            k_fill<<<(n + 255) / 256, 256>>>(inbox, n, SENT);
            void *args[] = {&inbox, &iters, &fence, &cyc, &res};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)k_allgather_push, dim3(sms), dim3(512), args, 0, 0);
            cudaDeviceSynchronize();
            long long h; double r; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&r, res, 8, cudaMemcpyDeviceToHost);
            printf("one-hop push all-gather (private inboxes), fence %d: %.0f cycles  [%s, total %.3f]\n", fence, (double)h / iters, cudaGetErrorString(e), r);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
