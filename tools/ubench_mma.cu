// Peak rates of the legacy (mma.sync) tensor-core paths on B200, register operands only: fp64 m8n8k4 (DMMA, the
// denominator of K2b's "% of DMMA peak"), TF32 m16n8k8 and BF16 m16n8k16 (the spectral GN preconditioner).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_mma.cu -o /tmp/ubench_mma
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int ILP>
__global__ void __launch_bounds__(256) k_mma(float *out, int iters)
{
    double d[ILP][2]; float c[ILP][4];
    for (int i = 0; i < ILP; i++) { d[i][0] = d[i][1] = 0.0; c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
    const double a64 = 1.0 + threadIdx.x * 1e-9, b64 = 1.0 - threadIdx.x * 1e-9;
    const unsigned int a32 = 0x3f800000u + threadIdx.x, b32 = 0x3f800000u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a64), "d"(b64));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a32), "r"(a32), "r"(a32), "r"(a32), "r"(b32), "r"(b32));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u));
        }
    }
    float s = 0.f;
    for (int i = 0; i < ILP; i++) s += (float)(d[i][0] + d[i][1]) + c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678f) out[0] = s;
}

template <int KIND>
void run(const char *name, double flop_per_instr, int sms)
{
    float *out; cudaMalloc(&out, 4);
    const int iters = 4096, ILP = 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps_per_sm : {4, 8, 16, 32}) {
        const int blocks = sms * warps_per_sm / 8;
        k_mma<KIND, ILP><<<blocks, 256>>>(out, 64);
        cudaEventRecord(e0);
        k_mma<KIND, ILP><<<blocks, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double instr = (double)blocks * 8 * iters * ILP;
        printf("%-28s %2d warps/SM: %8.2f TFLOP/s  (%.0f FLOP/clk/SM at 1965 MHz)\n", name, warps_per_sm, instr * flop_per_instr / ms / 1e9,
               instr * flop_per_instr / (ms * 1e-3) / sms / 1.965e9);
    }
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("DMMA m8n8k4 f64", 512.0, p.multiProcessorCount);
    run<1>("mma.sync m16n8k8 tf32", 2048.0, p.multiProcessorCount);
    run<2>("mma.sync m16n8k16 bf16", 4096.0, p.multiProcessorCount);
    return 0;
}
