// FP64 sin/cos issue-rate microbenchmark (roofline denominator of k4_jmn_kernel, cfg-5): how many full-range
// double-precision sin / cos / sincos evaluations per second one B200 sustains for arguments in [0, 630) (mu*beta at j <= 200),
// and the rate of the rotation recurrence (4 DFMA per step) that replaces most of them.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int iters, double x0, double *sink)
{
    double x = x0 + (blockIdx.x * blockDim.x + threadIdx.x) * 1e-3, acc = 0.0;
    double c = 1.0, s = 0.0; const double cd = cos(0.37), sd = sin(0.37);
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { double a, b; sincos(x, &a, &b); acc += a + b; }
        if (MODE == 1) acc += sin(x);
        if (MODE == 2) acc += cos(x);
        if (MODE == 3) { const double c2 = c * cd - s * sd; s = s * cd + c * sd; c = c2; acc += c * x; }
        x += 1.618;
        if (x > 630.0) x -= 630.0;
    }
    if (acc + c + s == 123.456) sink[0] = acc;
}
template <int MODE> void run(const char *name, double *sink)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(10, 0.1, sink);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(iters, 0.1, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-10s : %.3e evaluations/s\n", name, (double)blocks * threads * iters / (ms * 1e-3));
}
int main()
{
    double *sink; cudaMalloc(&sink, 8);
    run<0>("sincos", sink); run<1>("sin", sink); run<2>("cos", sink); run<3>("rotation", sink);
    return 0;
}
