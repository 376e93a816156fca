// Microbenchmark of the k2 main-loop DMMA pattern: 16 DMMAs per step, A operands fa[2], fb[2], B operands be[4], bo[4].
// Variants: (0) operands constant in registers, (1) operands reloaded from shared memory every step (prefetch one ahead),
// (2) as (1) plus the 4 DMULs of the A scaling.  Reports % of the 16-cycles-per-DMMA pipe rate.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(int iters, double *sink)
{
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc[2][4][2][2];
#pragma unroll
    for (int i = 0; i < 32; ++i) (&acc[0][0][0][0])[i] = lane * 1e-9;
    double fa[2], fb[2], be[4], bo[4], un = 1.0;
    const double *p = sm + lane;
    for (int i = 0; i < 2; ++i) { fa[i] = p[i * 32]; fb[i] = p[64 + i * 32]; }
    for (int g = 0; g < 4; ++g) { be[g] = p[128 + g * 64]; bo[g] = p[160 + g * 64]; }
    double nfa[2], nfb[2], nbe[4], nbo[4], nun = 1.0;
    for (int it = 0; it < iters; ++it) {
        if (MODE >= 1) {
            const double *q = sm + lane + ((it & 7) << 9);
            nun = q[420];
#pragma unroll
            for (int i = 0; i < 2; ++i) { nfa[i] = q[i * 32]; nfb[i] = q[64 + i * 32]; }
#pragma unroll
            for (int g = 0; g < 4; ++g) { nbe[g] = q[128 + g * 64]; nbo[g] = q[160 + g * 64]; }
        }
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < 2; ++i) { fa[i] *= un; fb[i] *= un; }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < 2; ++i) { dmma(acc[i][g][0], fa[i], be[g]); dmma(acc[i][g][1], fb[i], bo[g]); }
        if (MODE >= 1) {
            un = nun;
#pragma unroll
            for (int i = 0; i < 2; ++i) { fa[i] = nfa[i]; fb[i] = nfb[i]; }
#pragma unroll
            for (int g = 0; g < 4; ++g) { be[g] = nbe[g]; bo[g] = nbo[g]; }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += (&acc[0][0][0][0])[i];
    if (s == 123.456) sink[0] = s;
}
template <int MODE>
void run(int warps_per_sm, double *sink)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    k<MODE><<<148, warps_per_sm * 32>>>(50, sink);
    cudaEventRecord(e0);
    k<MODE><<<148, warps_per_sm * 32>>>(iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = (ms * 1e-3 * 1.965e9) / ((double)iters * 16 * warps_per_sm / 4.0);
    printf("mode %d warps/SM %2d : %.1f cycles per DMMA per sub-partition  (%.0f %% of the 16-cycle rate)\n", MODE, warps_per_sm, cyc, 1600.0 / cyc);
}
int main()
{
    double *sink; cudaMalloc(&sink, 8);
    for (int w : {4, 8, 12}) { run<0>(w, sink); run<1>(w, sink); run<2>(w, sink); }
    return 0;
}
