// Store-path microbenchmark, round 2: which ORDER of producing the cfg-2 output (10^5 column-major 201x201 matrices,
// 32.3 GB) lets one B200 write it fastest, restricted to orders the contraction kernel can actually produce
// (a warp owns a quadrant column qn for a chunk of angles; the output of a unit is column c = j+qn and its mirror
// N-1-c, for every angle of the chunk; with the beta <-> pi-beta pairing also the same two columns of the mirror
// matrices nb-1-b).
//   order 0  "parts"   : the round-1 drain: 64-row parts, per 256-byte piece all angles of the chunk (asc + desc target)
//   order 1  "column"  : whole 1608-byte columns, angle after angle (one warp writes a column front to back)
//   order 2  "column/sector-owned": like 1 but every warp writes the sector-aligned window
//             [ceil4(start), ceil4(end)) of its column: no 32-byte sector is ever written by two warps (the kernel
//             holds the <= 3 trailing elements back until the same warp produces the next column)
//   order 3  "column/sector-owned, TMA": like 2 through cp.async.bulk.global.shared::cta (one elected lane)
//   PAIR = 1 adds the mirror matrices as targets (4 columns per unit instead of 2).
//   static = 1: warp w owns a contiguous block of columns instead of taking them round-robin.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

struct Args { double *out; int N; long long nb; int A, order, pair, stat, piece_major; };

__device__ __forceinline__ void write_stretch(double *out, size_t e0, int len, int lane, bool owned)
{
    if (!owned) {
        for (int s = lane; s < len; s += 32) out[e0 + s] = 1.0;
    } else {
        const size_t a = (e0 + 3) & ~(size_t)3, b = (e0 + len + 3) & ~(size_t)3;
        for (size_t s = a + lane; s < b; s += 32) out[s] = 1.0;
    }
}

__global__ void __launch_bounds__(512) pattern(Args g)
{
    extern __shared__ __align__(128) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int N = g.N, A = g.A, Q = (N + 1) / 2, i0 = N - Q;
    const size_t NN = (size_t)N * N;
    const long long nb_eff = g.pair ? g.nb / 2 : g.nb;
    const long long nchunks = nb_eff / A;
    double *buf = sm + warp * 256;
    if (g.order == 3) { for (int i = lane; i < 256; i += 32) buf[i] = 1.0; __syncwarp(); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long b0 = ch * A;
        const int per = (Q + W - 1) / W;
        const int q_lo = g.stat ? warp * per : warp, q_hi = g.stat ? min(Q, q_lo + per) : Q, q_st = g.stat ? 1 : W;
        for (int qn = q_lo; qn < q_hi; qn += q_st) {
            const int cols[2] = {i0 + qn, N - 1 - i0 - qn};
            const int ncol = (qn >= (N & 1)) ? 2 : 1;
            for (int t = 0; t < ncol * (g.pair ? 2 : 1); ++t) {
                const int c = cols[t & 1];
                const bool mir = t >= 2;
                if (g.order == 0) {
                    // 64-row parts: asc rows i0+64p.. and desc rows; per piece of 32 rows, all angles
                    for (int p0 = 0; p0 < Q; p0 += 64) {
                        const int len = min(64, Q - p0);
                        for (int s = lane; s < len; s += 32)
                            for (int a = 0; a < A; ++a) {
                                const size_t m = (size_t)(mir ? g.nb - 1 - (b0 + a) : b0 + a) * NN + (size_t)c * N;
                                g.out[m + i0 + p0 + s] = 1.0;
                                if (p0 + s >= (N & 1)) g.out[m + (N - 1 - i0 - p0) - s] = 1.0;
                            }
                    }
                } else if (g.order == 3) {
                    for (int a = 0; a < A; ++a) {
                        const size_t e0 = (size_t)(mir ? g.nb - 1 - (b0 + a) : b0 + a) * NN + (size_t)c * N;
                        const size_t x = (e0 + 3) & ~(size_t)3, y = (e0 + N + 3) & ~(size_t)3;
                        if (lane == 0)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g.out + x), "r"(smem_addr(buf)), "r"((unsigned)((y - x) * 8)) : "memory");
                    }
                    if (lane == 0) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }
                    __syncwarp();
                } else if (g.piece_major) {
                    for (int s0 = 0; s0 < N + 3; s0 += 32)
                        for (int a = 0; a < A; ++a) {
                            const size_t e0 = (size_t)(mir ? g.nb - 1 - (b0 + a) : b0 + a) * NN + (size_t)c * N;
                            if (g.order == 1) { if (s0 + lane < N) g.out[e0 + s0 + lane] = 1.0; }
                            else { const size_t x = ((e0 + 3) & ~(size_t)3) + s0 + lane, y = (e0 + N + 3) & ~(size_t)3; if (x < y) g.out[x] = 1.0; }
                        }
                } else {
                    for (int a = 0; a < A; ++a) {
                        const size_t e0 = (size_t)(mir ? g.nb - 1 - (b0 + a) : b0 + a) * NN + (size_t)c * N;
                        write_stretch(g.out, e0, N, lane, g.order == 2);
                    }
                }
            }
        }
    }
    if (g.order == 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
    double *out; cudaMalloc(&out, (size_t)nb * 204 * 204 * 8 + 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const size_t n = (size_t)nb * 201 * 201;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); stream<<<148 * 8, 256>>>(out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("streaming write of %.2f GB: %.2f ms = %.0f GB/s\n", n * 8 / 1e9, ms, n * 8 / 1e6 / ms);
    cudaFuncSetAttribute(pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    auto run = [&](int N, int A, int W, int order, int pair, int stat, int pm) {
        Args g{out, N, nb, A, order, pair, stat, pm};
        const size_t nn = (size_t)nb * N * N;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); pattern<<<148, W * 32, W * 2048>>>(g); cudaEventRecord(e1);
            cudaError_t e = cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
        }
        printf("N %d angles/chunk %2d warps/SM %2d order %d pair %d static %d piece-major %d : %6.2f ms = %4.0f GB/s\n", N, A, W, order, pair,
               stat, pm, ms, nn * 8 / 1e6 / ms);
        fflush(stdout);
    };
    for (int N : {201, 200})
        for (int W : {4, 8, 16})
            for (int pair : {0, 1}) {
                run(N, 8, W, 0, pair, 0, 0);
                for (int order : {1, 2})
                    for (int stat : {0, 1})
                        for (int pm : {0, 1}) run(N, 8, W, order, pair, stat, pm);
                run(N, 8, W, 3, pair, 0, 0);
                run(N, 8, W, 3, pair, 1, 0);
            }
    for (int A : {4, 16}) for (int order : {1, 2}) for (int pair : {0, 1}) run(201, A, 8, order, pair, 1, 0);
    return 0;
}
