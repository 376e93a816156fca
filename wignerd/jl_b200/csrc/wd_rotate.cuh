// wd_rotate.cuh -- fused consumer of the d / D matrices: rotation of spherical-harmonic coefficients,
//     a'_{l m'} = sum_m D^l_{m m'}(alpha, beta, gamma) a_{l m}          (reference: docs/src/index.md:76-84)
// for a batch of rotations, WITHOUT ever forming a d matrix.  The Feng factorisation the whole engine rests on,
//     d^l_{m m'}(beta) = sum_mu <m|mu> e^{-i mu beta} <mu|m'>,   <m|mu> = (-i)^m u[m, mu]   (u: real eigenvectors of J_x),
// turns the rotation of one l into two real-matrix x complex-matrix products with diagonal phases in between:
//     c[m, b]  = (-i)^m e^{-i m alpha_b} a_{l m}(b)                      (k5_prephase_kernel)
//     z[mu, b] = e^{-i mu beta_b}  sum_m  u[m, mu] c[m, b]                (k5_gemm_kernel<true>,  A = u^T)
//     a'[m', b] = i^m' e^{-i m' gamma_b} sum_mu u[m', mu] z[mu, b]        (k5_gemm_kernel<false>, A = u)
// i.e. 8 N^2 flops per l and rotation instead of the ~N^3 / 2 of building d first (cfg-4, lmax = 512 on 4096 rotations:
// 5.9e12 instead of 1.4e14 flops, and 17 GB of coefficients instead of 5.9 TB of matrices).
//
// k5_gemm_kernel: FP64 tensor-core GEMM (mma.sync m8n8k4 -> DMMA.8x8x4), CTA tile 64 rows x 128 real columns (= 64
// complex right-hand sides; 64 columns when the batch is small), warps of 32 x 32, k chunks of 16 staged through shared memory with register prefetch of
// the next chunk; the phase factor of the epilogue is applied to the accumulators (a lane's two adjacent columns are
// the real and imaginary part of one coefficient).
#pragma once
#include "wd_k2_dmma.cuh"

namespace wd {

// full N x N eigenvector matrix of J_x from the quadrant table: U[i + ld * k] = u[m = i - l, mu = k - l] (column-major)
__global__ void __launch_bounds__(256) k5_expand_u_kernel(PlanDev p, double *__restrict__ U, int ld)
{
    const int N = p.N, K = p.Q, i0 = N - p.Q, k0 = N - K;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)N * N; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % N), k = (int)(e / N);
        // column k (mu = k - j): mu >= 0 is stored (kappa = k - k0), mu < 0 by u[i, -mu] = (-1)^i u[i, mu]
        const int kk = k >= k0 ? k : N - 1 - k;
        const int kappa = kk - k0;
        const int kp = wd_kp_of_kappa(kappa, K, p.K0p);
        const double eps = (kp < p.K0p) ? 1.0 : -1.0;                   // u[-m] = eps u[m]
        double v = (i >= i0) ? p.uq[(size_t)(i - i0) * p.ldu + kp] : eps * p.uq[(size_t)(N - 1 - i - i0) * p.ldu + kp];
        if (k < k0 && (i & 1)) v = -v;
        U[i + (size_t)ld * k] = v;
    }
}

// c[m][b] = (-i)^m e^{-i m alpha_b} a_{l m}: interleaved complex, row m has nb entries
__global__ void __launch_bounds__(256) k5_prephase_kernel(const double2 *__restrict__ alm, long long alm_stride, int l, const double *__restrict__ alpha,
                                                          long long nb, double2 *__restrict__ c)
{
    const int N = 2 * l + 1;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)N * nb; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e % nb;
        const int i = (int)(e / nb), m = i - l;
        const double2 a = alm[(size_t)b * alm_stride + (size_t)l * l + i];
        double s, cs;
        sincos(-(double)m * alpha[b], &s, &cs);
        // (-i)^m: m mod 4 = 0: 1, 1: -i, 2: -1, 3: i
        double pr = cs, pi = s;
        switch (m & 3) { case 1: { const double t = pr; pr = pi; pi = -t; } break; case 2: pr = -pr; pi = -pi; break; case 3: { const double t = pr; pr = -pi; pi = t; } break; default: break; }
        c[e] = make_double2(a.x * pr - a.y * pi, a.x * pi + a.y * pr);
    }
}

constexpr int K5_TM = 64, K5_TN = 128, K5_TK = 16;     // CTA tile: 64 rows x TN real columns (TN = 128, or 64 for small batches), k chunks of 16
constexpr int K5_LDA = K5_TK + 4;            // == 4 (mod 16): conflict-free A fragments

struct K5Args {
    const double *U;        // N x N column-major, leading dimension ld
    int ld, N, l;
    const double *in;       // [N][2 nb] (row k: nb interleaved complex numbers)
    double *out;            // TRANS_A: [N][2 nb] workspace;  else: user layout out[b * out_stride + l^2 + row] complex
    long long nb, out_stride;
    const double *angle;    // beta (TRANS_A) or gamma: epilogue phase e^{-i v(row) angle_b}, v = row - l
};

// C[row][col] = sum_k A[row][k] in[k][col];  TRANS_A: A[row][k] = U[k + ld * row] (z = u^T c), else A[row][k] = U[row + ld * k].
// TN real columns per CTA (2 * TN threads: 2 x TN/32 warps of 32 x 32): 128, or 64 when the batch is too small to fill the GPU.
template <bool TRANS_A, int TN>
__global__ void __launch_bounds__(2 * TN, TN == 64 ? 4 : 2) k5_gemm_kernel(K5Args a)
{
    constexpr int NT = 2 * TN, LDB = TN + 4;                         // LDB == 4 (mod 16): conflict-free B fragments
    __shared__ __align__(16) double sA[K5_TM * K5_LDA];
    __shared__ __align__(16) double sB[K5_TK * LDB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, c = lane & 3;
    const int N = a.N;
    const long long C = 2 * a.nb;
    const int row0 = blockIdx.y * K5_TM;
    const long long col0 = (long long)blockIdx.x * TN;
    const int wr = (warp / (TN / 32)) * 32, wc = (warp % (TN / 32)) * 32;           // warp tile origin inside the CTA tile

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 32; ++i) (&acc[0][0][0])[i] = 0.0;

    // global -> register prefetch of one k chunk: A 64 x 16 (1024 / NT per thread), B 16 x TN (8 per thread)
    constexpr int NA = 1024 / NT;
    double ra[NA], rb[8];
    auto gload = [&](const int k0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int e = tid + NT * i;
            int row, k;
            if (TRANS_A) { k = e & 15; row = e >> 4; }                 // k contiguous in memory
            else { row = e & 63; k = e >> 6; }                         // rows contiguous in memory
            const int gr = row0 + row, gk = k0 + k;
            ra[i] = (gr < N && gk < N) ? (TRANS_A ? a.U[gk + (size_t)a.ld * gr] : a.U[gr + (size_t)a.ld * gk]) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + NT * i, col = e % TN, k = e / TN;
            const int gk = k0 + k;
            const long long gc = col0 + col;
            rb[i] = (gk < N && gc < C) ? a.in[(size_t)gk * C + gc] : 0.0;
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int e = tid + NT * i;
            int row, k;
            if (TRANS_A) { k = e & 15; row = e >> 4; } else { row = e & 63; k = e >> 6; }
            sA[row * K5_LDA + k] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + NT * i, col = e % TN, k = e / TN;
            sB[k * LDB + col] = rb[i];
        }
    };
    gload(0);
    for (int k0 = 0; k0 < N; k0 += K5_TK) {
        __syncthreads();                                               // the previous chunk has been consumed
        sstore();
        __syncthreads();
        if (k0 + K5_TK < N) gload(k0 + K5_TK);                         // in flight during the DMMAs of this chunk
#pragma unroll
        for (int ks = 0; ks < K5_TK; ks += 4) {
            double fa[4], fb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) fa[i] = sA[(wr + 8 * i + r) * K5_LDA + ks + c];
#pragma unroll
            for (int jn = 0; jn < 4; ++jn) fb[jn] = sB[(ks + c) * LDB + wc + 8 * jn + r];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) dmma(acc[i][jn], fa[i], fb[jn]);
        }
    }
    // epilogue: the lane's two adjacent columns are (re, im) of rotation b = col / 2; multiply by the row's phase
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + wr + 8 * i + r;
        if (row >= N) continue;
        const int v = row - a.l;
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
            const long long col = col0 + wc + 8 * jn + 2 * c;
            if (col >= C) continue;
            const long long b = col >> 1;
            double s, cs;
            sincos(-(double)v * a.angle[b], &s, &cs);
            double pr = cs, pi = s;
            if (!TRANS_A) {                                            // i^m': 0: 1, 1: i, 2: -1, 3: -i
                switch (v & 3) { case 1: { const double t = pr; pr = -pi; pi = t; } break; case 2: pr = -pr; pi = -pi; break; case 3: { const double t = pr; pr = pi; pi = -t; } break; default: break; }
            }
            const double xr = acc[i][jn][0], xi = acc[i][jn][1];
            const double2 o = make_double2(xr * pr - xi * pi, xr * pi + xi * pr);
            if (TRANS_A) reinterpret_cast<double2 *>(a.out)[(size_t)row * a.nb + b] = o;
            else reinterpret_cast<double2 *>(a.out)[(size_t)b * a.out_stride + (size_t)a.l * a.l + row] = o;
        }
    }
}

}  // namespace wd
