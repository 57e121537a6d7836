// What fraction of the copy bandwidth can a kernel with the stream mix of K1 (6 reads + 1 write per cell) or K3 (4 reads +
// 6 writes) reach at all?  Plain grid-stride kernels over 33 M doubles per stream (the 1080x1920x16 volume), nothing but loads,
// one add per word and stores; 8-byte and 16-byte accesses.  The ceiling for the real kernels, which add halo reads and
// arithmetic on top.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_streams.cu -o /tmp/ubs
#include <cstdio>
#include <cuda_runtime.h>

template <int NR, int NW, typename T>
__global__ void __launch_bounds__(256) k_streams(size_t n, const T *__restrict__ in, T *__restrict__ out, size_t stride)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        T v[NR];
#pragma unroll
        for (int i = 0; i < NR; i++) v[i] = in[i * stride + k];
        T s = v[0];
#pragma unroll
        for (int i = 1; i < NR; i++) { if constexpr (sizeof(T) == 16) { s.x += v[i].x; s.y += v[i].y; } else s += v[i]; }
#pragma unroll
        for (int i = 0; i < NW; i++) out[i * stride + k] = s;
    }
}

// the access pattern of K1 / K3: a thread owns CPT consecutive cells of the (y, x) plane and marches through the planes
template <int NR, int NW, int CPT>
__global__ void __launch_bounds__(256) k_march(size_t P, int Nt, const double *__restrict__ in, double *__restrict__ out, size_t stride)
{
    for (size_t i0 = ((size_t)blockIdx.x * blockDim.x) * CPT; i0 < P; i0 += (size_t)gridDim.x * blockDim.x * CPT) {
        for (int n = 0; n < Nt; n++) {
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                const size_t i = i0 + (size_t)c * blockDim.x + threadIdx.x;      // CPT runs of 256 consecutive cells per block
                if (i >= P) continue;
                const size_t k = (size_t)n * P + i;
                double v[NR];
#pragma unroll
                for (int j = 0; j < NR; j++) v[j] = in[j * stride + k];
                double s = v[0];
#pragma unroll
                for (int j = 1; j < NR; j++) s += v[j];
#pragma unroll
                for (int j = 0; j < NW; j++) out[j * stride + k] = s;
            }
        }
    }
}

template <int NR, int NW, int CPT>
void run_march(const char *name, size_t P, int Nt, const double *in, double *out, int blocks)
{
    const size_t stride = P * Nt;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_march<NR, NW, CPT><<<blocks, 256>>>(P, Nt, in, out, stride);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-34s %2d reads + %d writes, marching in t, %d x 256 cells per block and plane, %5d blocks: %7.3f ms  %7.1f GB/s\n", name, NR, NW, CPT,
           blocks, best, (double)(NR + NW) * stride * 8 / best / 1e6);
}

template <int NR, int NW, typename T>
void run(const char *name, size_t cells, const double *in, double *out, int blocks)
{
    const size_t n = cells * 8 / sizeof(T), stride = n;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_streams<NR, NW, T><<<blocks, 256>>>(n, (const T *)in, (T *)out, stride);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-34s %2d reads + %d writes, %2zu-byte accesses, %5d blocks: %7.3f ms  %7.1f GB/s\n", name, NR, NW, sizeof(T), blocks, best,
           (double)(NR + NW) * cells * 8 / best / 1e6);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const size_t cells = (size_t)16 * 1080 * 1920;
    double *in, *out;
    cudaMalloc(&in, 6 * cells * 8); cudaMalloc(&out, 6 * cells * 8);
    cudaMemset(in, 0, 6 * cells * 8); cudaMemset(out, 0, 6 * cells * 8);
    for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
        run<1, 1, double>("copy", cells, in, out, blocks);
        run<1, 1, double2>("copy", cells, in, out, blocks);
        run<6, 1, double>("K1 mix", cells, in, out, blocks);
        run<6, 1, double2>("K1 mix", cells, in, out, blocks);
        run<4, 6, double>("K3 mix", cells, in, out, blocks);
        run<4, 6, double2>("K3 mix", cells, in, out, blocks);
        run<5, 5, double>("K2a streaming mix (one phase)", cells, in, out, blocks);
    }
    const size_t P = (size_t)1080 * 1920;
    for (int blocks : {148 * 8, 148 * 16, 8100}) {
        run_march<6, 1, 1>("K1 mix", P, 16, in, out, blocks);
        run_march<6, 1, 2>("K1 mix", P, 16, in, out, blocks);
        run_march<6, 1, 4>("K1 mix", P, 16, in, out, blocks);
        run_march<4, 6, 1>("K3 mix", P, 16, in, out, blocks);
        run_march<4, 6, 4>("K3 mix", P, 16, in, out, blocks);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
