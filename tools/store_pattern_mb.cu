// Store-path microbenchmark: how fast can one B200 write the cfg-2 output (10^5 matrices of NxN doubles, 32.3 GB at
// N = 201) for different orders in which the same bytes are produced?  Every warp instruction stores 256 contiguous
// bytes (the last one of a run is partial).
//   mode 0 ("tile"): a tile = RUN consecutive doubles of one column-major matrix for each of the A angles of a chunk;
//       a warp walks angle by angle; consecutive tile numbers are adjacent in memory; warps take tiles round-robin
//       so that the W warps of a CTA work on adjacent memory at the same time (the K2 pattern).
//   mode 1 ("matrix per warp"): warp w of the CTA writes whole matrices w, w+W, ... of the chunk front to back.
//   mode 2 ("tile, cta-split"): like mode 0 but S CTAs share one chunk (CTA s takes tiles = s mod S): fewer
//       matrices in flight on the whole GPU.
// N = 200 / 204 make every run 32-byte aligned (N = 201 shifts each column by 8 bytes): cost of partial sectors.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void pattern(double *out, int N, long long nb, int A, int RUN, int mode, int S)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const size_t NN = (size_t)N * N;
    const long long nchunks = nb / A;
    const int tiles = (int)((NN + RUN - 1) / RUN);          // tiles per matrix: the matrix cut into runs
    if (mode == 1) {
        for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
            for (int a = warp; a < A; a += W) {
                double *p = out + ((size_t)ch * A + a) * NN;
                for (size_t s = lane; s < NN; s += 32) p[s] = 1.0;
            }
        }
        return;
    }
    const long long nitems = nchunks * S;
    for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
        const long long ch = it / S;
        const int sp = (int)(it - ch * S);
        double *cb = out + (size_t)ch * A * NN;
        for (int t = warp + W * sp; t < tiles; t += W * S) {
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
    const long long nb = 100000;
    double *out; cudaMalloc(&out, (size_t)nb * 204 * 204 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const size_t n = (size_t)nb * 201 * 201;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); stream<<<148 * 8, 256>>>(out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("streaming write of %.2f GB: %.2f ms = %.0f GB/s\n", n * 8 / 1e9, ms, n * 8 / 1e6 / ms);
    auto run = [&](int N, int A, int RUN, int W, int mode, int S) {
        const size_t nn = (size_t)nb * N * N;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); pattern<<<148, W * 32>>>(out, N, nb, A, RUN, mode, S); cudaEventRecord(e1);
            cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("N %d mode %d angles/chunk %2d run %4d warps/SM %2d split %d : %6.2f ms = %4.0f GB/s\n", N, mode, A, RUN, W, S,
               ms, nn * 8 / 1e6 / ms);
    };
    for (int N : {201, 200, 204}) for (int RUN : {32, 64, N}) for (int W : {8, 16}) run(N, 16, RUN, W, 0, 1);
    for (int A : {4, 8, 32}) for (int RUN : {64, 201}) run(201, A, RUN, 8, 0, 1);
    for (int W : {8, 16}) run(201, 16, 0 + 64, W, 1, 1);
    for (int S : {2, 4, 8}) for (int RUN : {64, 201}) run(201, 16, RUN, 8, 2, S);
    return 0;
}
