// wd_kernels.cuh -- sm_100a device code of the Wigner-d/D engine.
//
// Mathematics (reference: jishnub/WignerD.jl src/WignerD.jl):
//   J_y = S J_x S^+, S = diag((-i)^m); J_x is the real symmetric tridiagonal with zero diagonal and
//   off-diagonals b_i = sqrt((i+1)(N-1-i))/2 (= X(j,n)/2 of :145,:153).  With u the real orthonormal
//   eigenvectors of J_x (eigenvalues mu = -j..j exactly) the reference's element sum (:187-206) is
//       d[i,i'] = sigma(i-i') * sum_{mu>=0} w_mu t(mu beta) u[i,mu] u[i',mu],
//   t = cos for even i-i', sin for odd; w_0 = 1, w_mu = 2; sigma(d) = -1 iff d mod 4 in {1,2}.
//   Only the quadrant m,n >= 0 of u is stored (UQ[q][kp]); the mu >= 0 eigenvectors are split into the
//   class that is symmetric under m -> -m (class 0) and the antisymmetric one (class 1), so that with
//   P0/P1 the partial sums over the two classes
//       d[ m, n] = s (P0+P1)   d[-m,-n] = s (P0+P1)   d[-m, n] = s (P0-P1)   d[ m,-n] = s (P0-P1)
//   (for half-integer j the mirrored pair takes the *other* trig function).
//
// Kernels:
//   k1_eig_kernel        batched eigensolver of J_x: twisted factorisation (the getvec step of MRRR,
//                        which is what LAPACK zheevr -- the reference's eigen!, :161 -- ends in) at the
//                        analytically known eigenvalues; one thread per eigenvector, all j in one launch.
//   k2_plain_kernel      FP64-ALU contraction, any j (fallback + on-GPU cross-check).
//   k2_dmma_kernel       the hot kernel: DMMA.8x8x4 contraction over (angle tile) x (m tile), UQ resident
//                        in shared memory, in-kernel trig tables, shared-memory staged coalesced stores,
//                        optional fused exp(-i(m alpha + n gamma)) epilogue (complex D).
//   k4_jmn_kernel        single elements (wignerdjmn / wignerDjmn): one warp per tuple.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wd {

// ------------------------------------------------------------------------------------------------
// per-j table ("plan") as seen by the device
// ------------------------------------------------------------------------------------------------
struct PlanDev {
    int twoj, N, Q;       // Q = K = (N+1)/2 : quadrant rows (m >= 0) and number of mu >= 0
    int K0, K1;           // class sizes (class 0: symmetric under m -> -m)
    int K0p, K1p, Kp;     // padded to multiples of 4; Kp = K0p + K1p
    int ldu;              // row stride of uq in doubles (Kp + 2, == 2 mod 4: conflict-free stride-2 rows)
    int zskip;            // 1 for integer j (m = 0 mirrors onto itself), else 0
    const double *uq;     // [Q][ldu], zero on the pads
    const double *mu;     // [Kp] eigenvalue of column kp (0 on pads)
    const double *w;      // [Kp] weight 1 (mu = 0) or 2, 0 on pads
};

__host__ __device__ inline int wd_round4(int x) { return (x + 3) & ~3; }

// kappa (0..K-1, mu = mu0 + kappa) -> padded column kp
__host__ __device__ inline int wd_kp_of_kappa(int kappa, int K, int K0p)
{
    const int c0 = (K - 1) & 1;                 // class 0: kappa == K-1 (mod 2)
    if (((kappa ^ c0) & 1) == 0) return (kappa - c0) >> 1;
    return K0p + ((kappa - (c0 ^ 1)) >> 1);
}

// ------------------------------------------------------------------------------------------------
// K1: batched tridiagonal eigensolver (twisted factorisation at the known eigenvalues)
// ------------------------------------------------------------------------------------------------
struct EigJob {
    int twoj, N, K, K0p, ldu;
    double *uq;        // out [Q][ldu]
    double *mu, *w;    // out [Kp]
    double *scratch;   // 2*N*K doubles
    int Kp;
};

__global__ void __launch_bounds__(128) k1_eig_kernel(const EigJob *__restrict__ jobs)
{
    const EigJob jb = jobs[blockIdx.y];
    const int kappa = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = jb.N, K = jb.K;
    // pads of mu / w (thread-strided, done by block 0)
    if (blockIdx.x == 0) {
        for (int kp = threadIdx.x; kp < jb.Kp; kp += blockDim.x) {
            bool used;
            if (kp < jb.K0p) used = kp < (K + 1) / 2; else used = (kp - jb.K0p) < K / 2;
            if (!used) { jb.mu[kp] = 0.0; jb.w[kp] = 0.0; }
        }
    }
    if (kappa >= K) return;
    const double lam = 0.5 * (double)(2 * kappa + (jb.twoj & 1));
    double *Dp = jb.scratch + kappa;                       // Dp[i*K]
    double *Z = jb.scratch + (size_t)N * K + kappa;        // Z[i*K]
    const double pivmin = 1e-200;

    // stationary qd (top-down LDL^T of T - lam I); by persymmetry of T the bottom-up pivots are D-[i] = D+[N-1-i]
    double d = -lam;
    if (fabs(d) < pivmin) d = -pivmin;
    Dp[0] = d;
    for (int i = 0; i < N - 1; ++i) {
        const double b2 = 0.25 * ((double)(i + 1) * (double)(N - 1 - i));
        d = -lam - b2 / d;
        if (fabs(d) < pivmin) d = -pivmin;
        Dp[(size_t)(i + 1) * K] = d;
    }
    // twist index: argmin |gamma_r|, gamma_r = D+[r] + D-[r] + lam
    int r = 0;
    double best = 1.7976931348623157e308;
    for (int i = 0; i < N; ++i) {
        const double g = fabs(Dp[(size_t)i * K] + Dp[(size_t)(N - 1 - i) * K] + lam);
        if (g < best) { best = g; r = i; }
    }
    double nrm = 1.0, z = 1.0;
    Z[(size_t)r * K] = 1.0;
    for (int i = r - 1; i >= 0; --i) {
        const double b = 0.5 * sqrt((double)(i + 1) * (double)(N - 1 - i));
        z = -(b / Dp[(size_t)i * K]) * z;
        Z[(size_t)i * K] = z;
        nrm += z * z;
    }
    z = 1.0;
    for (int i = r + 1; i < N; ++i) {
        const double b = 0.5 * sqrt((double)i * (double)(N - i));
        z = -(b / Dp[(size_t)(N - 1 - i) * K]) * z;
        Z[(size_t)i * K] = z;
        nrm += z * z;
    }
    double inv = 1.0 / sqrt(nrm);
    if (Z[(size_t)(N - 1) * K] < 0.0) inv = -inv;          // deterministic sign: u[m=+j, mu] > 0
    const int Q = K, i0 = N - Q;
    const int kp = wd_kp_of_kappa(kappa, K, jb.K0p);
    for (int q = 0; q < Q; ++q) jb.uq[(size_t)q * jb.ldu + kp] = Z[(size_t)(i0 + q) * K] * inv;
    jb.mu[kp] = lam;
    jb.w[kp] = (lam == 0.0) ? 1.0 : 2.0;
}

// sigma(di): -1 iff di mod 4 in {1,2}; returned as the sign-bit mask for the high word
__device__ __forceinline__ unsigned sigma_bit(int di) { return ((0x6u >> (di & 3)) & 1u) << 31; }
__device__ __forceinline__ double flip(double v, unsigned signbit)
{
    return __hiloint2double(__double2hiint(v) ^ (int)signbit, __double2loint(v));
}

// ------------------------------------------------------------------------------------------------
// K2 (plain): one thread per quadrant pair and angle.  Any j.  out real or complex.
// ------------------------------------------------------------------------------------------------
struct K2Args {
    PlanDev p;
    const double *beta, *alpha, *gamma;   // alpha/gamma only for complex output
    long long nb;
    double *out;                          // nb * N*N doubles (or * 2 for complex)
    int equator;                          // apply the exact zeros of :230-237
};

template <bool CPLX>
__device__ __forceinline__ void emit(const K2Args &a, long long b, int row, int col, double v,
                                     double alpha, double gamma)
{
    const PlanDev &p = a.p;
    if (a.equator && (p.twoj & 1) == 0) {
        const int j = p.twoj >> 1;                       // row = j+m, col = j+n
        if ((col == j && (row & 1)) || (row == j && (col & 1))) v = 0.0;   // (j+m odd && n==0) || (j+n odd && m==0)
    }
    const size_t e = ((size_t)b * p.N + col) * p.N + row;
    if (!CPLX) {
        a.out[e] = v;
    } else {
        const double m = 0.5 * (double)(2 * row - p.twoj), n = 0.5 * (double)(2 * col - p.twoj);
        double s, c;
        sincos(-(m * alpha + n * gamma), &s, &c);        // exactly :339 / :245
        reinterpret_cast<double2 *>(a.out)[e] = make_double2(v * c, v * s);
    }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) k2_plain_kernel(K2Args a)
{
    extern __shared__ double sm[];
    const PlanDev &p = a.p;
    double *tc = sm, *ts = sm + p.Kp;
    const long long b = blockIdx.x;
    const double beta = a.beta[b];
    for (int kp = threadIdx.x; kp < p.Kp; kp += blockDim.x) {
        double s, c;
        sincos(p.mu[kp] * beta, &s, &c);
        const double w = p.w[kp];
        tc[kp] = w * c;
        ts[kp] = w * s;
    }
    __syncthreads();
    const double alpha = CPLX ? a.alpha[b] : 0.0, gamma = CPLX ? a.gamma[b] : 0.0;
    const int Q = p.Q, N = p.N, i0 = N - Q;
    const bool half = p.twoj & 1;
    for (int pair = blockIdx.y * blockDim.x + threadIdx.x; pair < Q * Q; pair += gridDim.y * blockDim.x) {
        const int qm = pair % Q, qn = pair / Q;
        const double *um = p.uq + (size_t)qm * p.ldu, *un = p.uq + (size_t)qn * p.ldu;
        const bool odd = (qm - qn) & 1;
        const double *t_main = odd ? ts : tc, *t_alt = odd ? tc : ts;
        double P0 = 0, P1 = 0, A0 = 0, A1 = 0;
        for (int kp = 0; kp < p.K0p; ++kp) {
            const double g = um[kp] * un[kp];
            P0 = fma(t_main[kp], g, P0);
            if (half) A0 = fma(t_alt[kp], g, A0);
        }
        for (int kp = p.K0p; kp < p.Kp; ++kp) {
            const double g = um[kp] * un[kp];
            P1 = fma(t_main[kp], g, P1);
            if (half) A1 = fma(t_alt[kp], g, A1);
        }
        if (!half) { A0 = P0; A1 = P1; }
        const int i = i0 + qm, ip = i0 + qn, ir = N - 1 - i, ipr = N - 1 - ip;
        const double sum = P0 + P1, dif = A0 - A1;
        emit<CPLX>(a, b, i, ip, flip(sum, sigma_bit(i - ip)), alpha, gamma);
        if (qm >= p.zskip && qn >= p.zskip) emit<CPLX>(a, b, ir, ipr, flip(sum, sigma_bit(ir - ipr)), alpha, gamma);
        if (qm >= p.zskip) emit<CPLX>(a, b, ir, ip, flip(dif, sigma_bit(ir - ip)), alpha, gamma);
        if (qn >= p.zskip) emit<CPLX>(a, b, i, ipr, flip(dif, sigma_bit(i - ipr)), alpha, gamma);
    }
}

// ------------------------------------------------------------------------------------------------
// K2 (DMMA): the hot kernel.
//   CTA = 8 warps, persistent over chunks of 16 angles.  Shared memory: UQ (all of it), the mu / w
//   vectors, the chunk's trig tables Tc/Ts[16][ldt] (w cos(mu beta), w sin(mu beta)), optional phase
//   tables, and one staging tile per warp.
//   Warp unit = one quadrant column qn  x  up to GMAX groups of 16 quadrant rows  x  16 angles.
//   MMA roles (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4):  A[8 angles][4 kp] = T(beta, kp) * UQ[qn][kp],
//   B[4 kp][8 rows] = UQ[row][kp], C[8 angles][8 rows].  A 16-row group is two tiles: the even rows and
//   the odd rows (so one tile needs one trig function); a lane then owns 4 consecutive rows.
// ------------------------------------------------------------------------------------------------
constexpr int K2_WARPS = 8;
constexpr int K2_THREADS = K2_WARPS * 32;
constexpr int K2_ANG = 16;                 // angles per chunk (2 MMA row tiles)
constexpr int K2_STG_LD = 66;              // staging row stride (doubles): == 2 mod 4, 16B-aligned rows
constexpr int K2_STG_ROWS = 8;

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d[0]), "+d"(d[1])
        : "d"(a), "d"(b));
}

struct K2Smem {                 // offsets in doubles, computed identically on host and device
    int uq, mu, w, tc, ts, ph, stg, total;
    int ldt, qpad;
};
__host__ __device__ inline K2Smem k2_smem_layout(int Q, int Kp, int ldu, bool cplx)
{
    K2Smem s;
    s.ldt = Kp | 4;                        // == 4 mod 8: conflict-free for 8 consecutive rows x 4 cols
    s.qpad = (Q + 15) & ~15;               // UQ rows padded with zeros to whole 16-row groups
    int o = 0;
    s.uq = o;  o += s.qpad * ldu;
    s.mu = o;  o += Kp;
    s.w = o;   o += Kp;
    s.tc = o;  o += K2_ANG * s.ldt;
    s.ts = o;  o += K2_ANG * s.ldt;
    s.ph = o;  o += cplx ? 4 * K2_ANG * Q : 0;          // cos(m a), sin(m a), cos(n g), sin(n g)  [16][Q]
    o = (o + 1) & ~1;
    s.stg = o; o += K2_WARPS * K2_STG_ROWS * K2_STG_LD;
    s.total = o;
    return s;
}

struct K2Ctx {                  // per-thread constants of the DMMA kernel
    const double *sUQ, *sTc, *sTs, *sPh;
    double *stg;
    int ldu, ldt, N, Q, i0, zskip, Kp, K0p, twoj;
    int lane, r, c;
    long long b0;
};

// One warp unit: quadrant column qn, GC groups of 16 quadrant rows starting at group g0, 16 angles.
template <bool HALF, bool CPLX, int GC>
__device__ __forceinline__ void k2_unit(const K2Args &a, const K2Ctx &cx, const int qn, const int g0)
{
    constexpr int NALT = HALF ? 2 : 1;
    const int ldu = cx.ldu, ldt = cx.ldt, r = cx.r, c = cx.c;
    const int pe = qn & 1;
    // even-row tiles pair with cos when qn is even, with sin when qn is odd
    const double *tA = (pe ? cx.sTs : cx.sTc) + r * ldt + c;     // trig for the even-row tiles
    const double *tB = (pe ? cx.sTc : cx.sTs) + r * ldt + c;     // trig for the odd-row tiles
    const double *un = cx.sUQ + qn * ldu + c;
    const double *ub = cx.sUQ + (16 * g0 + 2 * r) * ldu + c;     // even rows; odd rows at +ldu; groups at +16*ldu
    const int gs = 16 * ldu, t8 = 8 * ldt;

    double acc[2][2][GC][2][NALT][2];            // [class][angle tile][group][even/odd rows][main/alt][frag]
#pragma unroll
    for (int i = 0; i < 2 * 2 * GC * 2 * NALT * 2; ++i) (&acc[0][0][0][0][0][0])[i] = 0.0;

#pragma unroll
    for (int cls = 0; cls < 2; ++cls) {
        const int kbeg = cls ? cx.K0p : 0;
        const int kend = cls ? cx.Kp : cx.K0p;
        // software pipeline: fragments of step k+4 are loaded while the DMMAs of step k issue.  The
        // prefetch of the step after the last one reads (and discards) in-bounds shared memory.
        double n_un = un[kbeg], n_ta[2], n_tb[2], n_be[GC], n_bo[GC];
#pragma unroll
        for (int ai = 0; ai < 2; ++ai) { n_ta[ai] = tA[ai * t8 + kbeg]; n_tb[ai] = tB[ai * t8 + kbeg]; }
#pragma unroll
        for (int g = 0; g < GC; ++g) { n_be[g] = ub[g * gs + kbeg]; n_bo[g] = ub[g * gs + ldu + kbeg]; }
#pragma unroll 1
        for (int k = kbeg; k < kend; k += 4) {
            double fa[2], fb[2], be[GC], bo[GC];
#pragma unroll
            for (int ai = 0; ai < 2; ++ai) { fa[ai] = n_ta[ai] * n_un; fb[ai] = n_tb[ai] * n_un; }
#pragma unroll
            for (int g = 0; g < GC; ++g) { be[g] = n_be[g]; bo[g] = n_bo[g]; }
            const int kn = k + 4;
            n_un = un[kn];
#pragma unroll
            for (int ai = 0; ai < 2; ++ai) { n_ta[ai] = tA[ai * t8 + kn]; n_tb[ai] = tB[ai * t8 + kn]; }
#pragma unroll
            for (int g = 0; g < GC; ++g) { n_be[g] = ub[g * gs + kn]; n_bo[g] = ub[g * gs + ldu + kn]; }
#pragma unroll
            for (int g = 0; g < GC; ++g) {
#pragma unroll
                for (int ai = 0; ai < 2; ++ai) {
                    dmma(acc[cls][ai][g][0][0], fa[ai], be[g]);
                    dmma(acc[cls][ai][g][1][0], fb[ai], bo[g]);
                    if (HALF) {
                        dmma(acc[cls][ai][g][0][1], fb[ai], be[g]);
                        dmma(acc[cls][ai][g][1][1], fa[ai], bo[g]);
                    }
                }
            }
        }
    }

    // ---- epilogue: butterfly, signs, stage through shared memory, coalesced stores ------------------
    const int N = cx.N, Q = cx.Q, i0 = cx.i0, zskip = cx.zskip, lane = cx.lane;
    double *stg = cx.stg;
    const int ip = i0 + qn, ipr = N - 1 - ip;
    const int qbase = 16 * g0;                     // first quadrant row of the unit
    constexpr int span = 16 * GC;                  // rows covered (<= 64)
#pragma unroll
    for (int tgt = 0; tgt < 4; ++tgt) {
        // tgt 0: ( i, i')  sum   ascending      tgt 1: (ir, i'r) sum  descending
        // tgt 2: (ir, i')  dif   descending     tgt 3: ( i, i'r) dif  ascending
        const bool desc = (tgt == 1 || tgt == 2);
        const bool dif = (tgt >= 2);
        const int col = (tgt == 0 || tgt == 2) ? ip : ipr;
        if ((tgt == 1 || tgt == 3) && qn < zskip) continue;         // n = 0 mirrors onto itself (warp-uniform)
        // global row of staging position s: ascending i0+qbase+s ; descending (N-1-i0-qbase-(span-1)) + s
        const int row0 = desc ? (N - 1 - i0 - qbase - (span - 1)) : (i0 + qbase);
        // sign: di of the lane's element t is +-(i0+qbase+16g+4c+t) ... only di mod 4 matters
        const int dib = desc ? (N - 1 - i0 - qbase - col) : (i0 + qbase - col);
#pragma unroll
        for (int ai = 0; ai < 2; ++ai) {
            __syncwarp();
#pragma unroll
            for (int g = 0; g < GC; ++g) {
                double v[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int eo = t & 1, h = t >> 1;
                    constexpr int dummy = 0; (void)dummy;
                    const int alt = (HALF && dif) ? 1 : 0;
                    const double p0 = acc[0][ai][g][eo][alt][h], p1 = acc[1][ai][g][eo][alt][h];
                    const double x = dif ? (p0 - p1) : (p0 + p1);
                    v[t] = flip(x, sigma_bit(desc ? (dib - t) : (dib + t)));
                }
                if (!desc) {
                    double2 *dst = reinterpret_cast<double2 *>(stg + r * K2_STG_LD + 16 * g + 4 * c);
                    dst[0] = make_double2(v[0], v[1]);
                    dst[1] = make_double2(v[2], v[3]);
                } else {
                    double2 *dst = reinterpret_cast<double2 *>(stg + r * K2_STG_LD + (span - 4) - 16 * g - 4 * c);
                    dst[0] = make_double2(v[3], v[2]);
                    dst[1] = make_double2(v[1], v[0]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int ar = 0; ar < 8; ++ar) {
                const long long b = cx.b0 + ai * 8 + ar;
                if (b >= a.nb) break;              // warp-uniform
#pragma unroll
                for (int hh = 0; hh < (span + 31) / 32; ++hh) {
                    const int s = hh * 32 + lane;
                    if (s >= span) continue;
                    const int qm = desc ? (qbase + span - 1 - s) : (qbase + s);
                    if (qm >= Q || (desc && qm < zskip)) continue;
                    const int row = row0 + s;
                    double v = stg[ar * K2_STG_LD + s];
                    if (a.equator && zskip) {
                        const int j = cx.twoj >> 1;
                        if ((col == j && (row & 1)) || (row == j && (col & 1))) v = 0.0;
                    }
                    const size_t e = ((size_t)b * N + col) * N + row;
                    if (!CPLX) {
                        a.out[e] = v;
                    } else {
                        // exp(-i(m alpha + n gamma)) = (ca - i sa)(cg - i sg); m -> -m flips sa, n -> -n flips sg
                        const int ang = ai * 8 + ar;
                        const bool mneg = row < i0;
                        const int qr = mneg ? (N - 1 - row - i0) : (row - i0);
                        const int qc = (col >= i0) ? (col - i0) : (N - 1 - col - i0);
                        const double ca = cx.sPh[(0 * K2_ANG + ang) * Q + qr];
                        double sa = cx.sPh[(1 * K2_ANG + ang) * Q + qr];
                        const double cg = cx.sPh[(2 * K2_ANG + ang) * Q + qc];
                        double sg = cx.sPh[(3 * K2_ANG + ang) * Q + qc];
                        if (mneg) sa = -sa;
                        if (col < i0) sg = -sg;
                        const double cr = ca * cg - sa * sg;
                        const double ci = -(sa * cg + ca * sg);
                        reinterpret_cast<double2 *>(a.out)[e] = make_double2(v * cr, v * ci);
                    }
                }
            }
        }
    }
}

template <bool HALF, bool CPLX>
__global__ void __launch_bounds__(K2_THREADS, 1) k2_dmma_kernel(K2Args a)
{
    constexpr int GMAX = HALF ? 2 : 4;
    extern __shared__ __align__(16) double sm[];
    const PlanDev &p = a.p;
    const int Q = p.Q, Kp = p.Kp, ldu = p.ldu;
    const K2Smem L = k2_smem_layout(Q, Kp, ldu, CPLX);
    const int ldt = L.ldt;
    double *sUQ = sm + L.uq, *sMu = sm + L.mu, *sW = sm + L.w, *sTc = sm + L.tc, *sTs = sm + L.ts;
    double *sPh = sm + L.ph;
    const int tid = threadIdx.x, warp = tid >> 5;
    K2Ctx cx;
    cx.sUQ = sUQ; cx.sTc = sTc; cx.sTs = sTs; cx.sPh = sPh;
    cx.stg = sm + L.stg + warp * (K2_STG_ROWS * K2_STG_LD);
    cx.ldu = ldu; cx.ldt = ldt; cx.N = p.N; cx.Q = Q; cx.i0 = p.N - Q; cx.zskip = p.zskip;
    cx.Kp = Kp; cx.K0p = p.K0p; cx.twoj = p.twoj;
    cx.lane = tid & 31; cx.r = cx.lane >> 2; cx.c = cx.lane & 3;

    // one-time: UQ (+ zero pad rows), mu, w -> shared (16-byte vector copies; ldu and Kp are even)
    {
        const double2 *src = reinterpret_cast<const double2 *>(p.uq);
        double2 *dst = reinterpret_cast<double2 *>(sUQ);
        const int nvalid = (Q * ldu) / 2, nall = (L.qpad * ldu) / 2;
        for (int i = tid; i < nall; i += K2_THREADS) dst[i] = (i < nvalid) ? src[i] : make_double2(0.0, 0.0);
        for (int i = tid; i < Kp; i += K2_THREADS) { sMu[i] = p.mu[i]; sW[i] = p.w[i]; }
    }
    const int ngroups = L.qpad >> 4;
    const int upc = (ngroups + GMAX - 1) / GMAX;          // units per column
    const int nunits = Q * upc;
    const long long nchunks = (a.nb + K2_ANG - 1) / K2_ANG;

    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long b0 = ch * K2_ANG;
        cx.b0 = b0;
        __syncthreads();                                   // previous chunk's readers are done (and UQ is in)
        for (int idx = tid; idx < K2_ANG * Kp; idx += K2_THREADS) {
            const int ang = idx / Kp, kp = idx - ang * Kp;
            long long b = b0 + ang; if (b >= a.nb) b = a.nb - 1;
            double s, cs;
            sincos(sMu[kp] * a.beta[b], &s, &cs);         // fl(mu*beta) then sincos, like :199
            const double w = sW[kp];
            sTc[ang * ldt + kp] = w * cs;
            sTs[ang * ldt + kp] = w * s;
        }
        if (CPLX) {
            for (int idx = tid; idx < K2_ANG * Q; idx += K2_THREADS) {
                const int ang = idx / Q, q = idx - ang * Q;
                long long b = b0 + ang; if (b >= a.nb) b = a.nb - 1;
                const double m = 0.5 * (double)(2 * q + (p.twoj & 1));
                double s, cs;
                sincos(m * a.alpha[b], &s, &cs);
                sPh[(0 * K2_ANG + ang) * Q + q] = cs; sPh[(1 * K2_ANG + ang) * Q + q] = s;
                sincos(m * a.gamma[b], &s, &cs);
                sPh[(2 * K2_ANG + ang) * Q + q] = cs; sPh[(3 * K2_ANG + ang) * Q + q] = s;
            }
        }
        __syncthreads();

        for (int u = warp; u < nunits; u += K2_WARPS) {
            const int qn = u / upc;
            const int g0 = (u - qn * upc) * GMAX;
            const int gcnt = min(GMAX, ngroups - g0);
            if (HALF) {
                if (gcnt == 2) k2_unit<HALF, CPLX, 2>(a, cx, qn, g0);
                else k2_unit<HALF, CPLX, 1>(a, cx, qn, g0);
            } else {
                switch (gcnt) {
                case 4: k2_unit<HALF, CPLX, 4>(a, cx, qn, g0); break;
                case 3: k2_unit<HALF, CPLX, 3>(a, cx, qn, g0); break;
                case 2: k2_unit<HALF, CPLX, HALF ? 1 : 2>(a, cx, qn, g0); break;
                default: k2_unit<HALF, CPLX, 1>(a, cx, qn, g0); break;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4: single elements.  One warp per tuple; lanes stride the padded mu axis.
// ------------------------------------------------------------------------------------------------
struct K4Args {
    const PlanDev *plans;     // indexed by twoj (entries with uq == nullptr are unprepared)
    int max_twoj;
    const int32_t *twoj, *twom, *twon;
    const double *beta, *alpha, *gamma;
    long long n;
    int beta_kind;            // wd_beta_kind
    double *out;
    int *err;                 // set to 1 on an invalid tuple
};

template <bool CPLX>
__global__ void __launch_bounds__(256) k4_jmn_kernel(K4Args a)
{
    const int lane = threadIdx.x & 31;
    const long long wpg = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < a.n; t += wpg) {
        const int tj = a.twoj[t], tm = a.twom[t], tn = a.twon[t];
        const int kind = a.beta_kind;
        double val = 0.0;
        bool done = false;
        // special co-latitudes (src/WignerD.jl:218-237); the poles do no index checks, like the reference
        if (kind == 1) { val = (tm == tn) ? 1.0 : 0.0; done = true; }
        else if (kind == 2) {
            val = (tm == -tn) ? ((((tj - tn) >> 1) & 1) ? -1.0 : 1.0) : 0.0; done = true;
        } else if (kind == 3 && ((tj + tm) & 1) == 0 && ((tj + tn) & 1) == 0) {
            const bool jm_odd = ((tj + tm) >> 1) & 1, jn_odd = ((tj + tn) >> 1) & 1;
            if ((jm_odd && tn == 0) || (jn_odd && tm == 0)) { val = 0.0; done = true; }
        }
        if (!done) {
            const bool ok = tj >= 0 && tj <= a.max_twoj && ((tj + tm) & 1) == 0 && ((tj + tn) & 1) == 0 &&
                            tm >= -tj && tm <= tj && tn >= -tj && tn <= tj;
            if (!ok) { if (lane == 0) *a.err = 1; continue; }
            const PlanDev p = a.plans[tj];
            const double beta = (kind == 3) ? 1.5707963267948966 : a.beta[t];
            const int N = p.N, i0 = N - p.Q;
            const int i = (tm + tj) >> 1, ip = (tn + tj) >> 1;
            const int qm = (i >= i0) ? i - i0 : (N - 1 - i) - i0;
            const int qn = (ip >= i0) ? ip - i0 : (N - 1 - ip) - i0;
            const bool flip1 = ((i < i0) != (ip < i0));       // one mirrored index: class 1 changes sign
            const bool odd = (i - ip) & 1;
            const double *um = p.uq + (size_t)qm * p.ldu, *un = p.uq + (size_t)qn * p.ldu;
            double s0 = 0.0, s1 = 0.0;
            for (int kp = lane; kp < p.Kp; kp += 32) {
                double s, c;
                sincos(p.mu[kp] * beta, &s, &c);
                const double g = p.w[kp] * (odd ? s : c) * (um[kp] * un[kp]);
                if (kp < p.K0p) s0 += g; else s1 += g;
            }
            double x = flip1 ? (s0 - s1) : (s0 + s1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            val = flip(x, sigma_bit(i - ip));
        }
        if (lane == 0) {
            if (!CPLX) a.out[t] = val;
            else {
                const double m = 0.5 * (double)tm, n = 0.5 * (double)tn;
                double s, c;
                sincos(-(a.alpha[t] * m + a.gamma[t] * n), &s, &c);      // :245
                reinterpret_cast<double2 *>(a.out)[t] = make_double2(val * c, val * s);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// small helpers: pole fills (src/WignerD.jl:247-264) on device buffers, FP64 peak micro-benchmarks
// ------------------------------------------------------------------------------------------------
__global__ void fp64_peak_kernel(int kind, int iters, double *sink)
{
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
    const double x = 1.0 + threadIdx.x * 1e-12, y = 1.0 - threadIdx.x * 1e-12;
    if (kind == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(acc[i], x, y);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[i][0] = fma(acc[i][0], x, y); acc[i][1] = fma(acc[i][1], y, x); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) sink[0] = s;
}

}  // namespace wd
