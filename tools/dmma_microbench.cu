// DMMA.8x8x4 issue-rate microbenchmark: TFLOP/s vs warps per SM and independent accumulator chains per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_mb tools/dmma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(int iters, double *sink)
{
    double acc[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
    const double x = 1.0 + threadIdx.x * 1e-12, y = 1.0 - threadIdx.x * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(acc[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) sink[0] = s;
}
template <int ILP>
void run(int warps_per_sm, double *sink)
{
    int dev_sms = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000 / ILP * 4;
    k<ILP><<<dev_sms, warps_per_sm * 32>>>(100, sink);
    cudaEventRecord(e0);
    k<ILP><<<dev_sms, warps_per_sm * 32>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)dev_sms * warps_per_sm * iters * ILP * 512.0;
    double cyc_per_dmma_per_smsp = (ms * 1e-3 * 1.965e9) / ((double)iters * ILP * warps_per_sm / 4.0);
    printf("warps/SM %2d  ILP %2d : %7.2f TFLOP/s   ~%.1f cycles per DMMA per SMSP (at 1965 MHz)\n", warps_per_sm, ILP,
           flops / (ms * 1e-3) / 1e12, cyc_per_dmma_per_smsp);
}
int main()
{
    double *sink; cudaMalloc(&sink, 8);
    for (int w : {4, 8, 16, 32}) { run<1>(w, sink); run<2>(w, sink); run<4>(w, sink); run<8>(w, sink); run<16>(w, sink); run<32>(w, sink); }
    return 0;
}
