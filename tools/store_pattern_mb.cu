// Store-path microbenchmark: how fast can one B200 write the cfg-2 output (10^5 matrices of 201x201 doubles, 32.3 GB)
// for different orders in which the same bytes are produced?  Every warp instruction stores 256 contiguous bytes
// (the last one of a run is partial).  A "tile" = RUN consecutive doubles of one column for each of the A angles of
// a chunk; a warp walks angle by angle.  Tiles of a chunk are numbered so that consecutive tile numbers are
// adjacent in memory (column-major); warps take tiles round-robin (w, w+W, ...) so that the W warps of a CTA work
// on adjacent memory at the same time.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void pattern(double *out, int N, long long nb, int A, int RUN)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const size_t NN = (size_t)N * N;
    const long long nchunks = nb / A;
    const int tiles = (int)((NN + RUN - 1) / RUN);          // tiles per chunk: the matrix cut into runs
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        double *cb = out + (size_t)ch * A * NN;
        for (int t = warp; t < tiles; t += W) {
            const size_t e = (size_t)t * RUN;
            const int len = (int)min((size_t)RUN, NN - e);
            for (int a = 0; a < A; ++a) {
                double *p = cb + (size_t)a * NN + e;
                for (int s = lane; s < len; s += 32) p[s] = 1.0;
            }
        }
    }
}
__global__ void stream(double *out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = 1.0;
}
int main()
{
    const int N = 201; const long long nb = 100000;
    const size_t n = (size_t)nb * N * N;
    double *out; cudaMalloc(&out, n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); stream<<<148 * 8, 256>>>(out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("streaming write of %.2f GB: %.2f ms = %.0f GB/s\n", n * 8 / 1e9, ms, n * 8 / 1e6 / ms);
    const int As[] = {16}; const int RUNs[] = {32, 64, 128, 201}; const int Ws[] = {8, 16, 32};
    for (int A : As) for (int RUN : RUNs) for (int W : Ws) {
        cudaEventRecord(e0); pattern<<<148, W * 32>>>(out, N, nb, A, RUN); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("angles/chunk %2d  run %4d doubles  warps/SM %d : %6.2f ms = %4.0f GB/s\n", A, RUN, W, ms, n * 8 / 1e6 / ms);
    }
    return 0;
}
