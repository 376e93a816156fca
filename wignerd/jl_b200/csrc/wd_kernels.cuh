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
//   k2_real_kernel       (wd_k2_dmma.cuh) the hot kernel: DMMA.8x8x4 contraction over (angle tile) x (m tile), UQ resident
//                        in shared memory, trig tables made one chunk ahead, store warps writing two mirror targets
//                        per staged tile; k2_cplx_kernel: the same with a fused exp(-i(m alpha + n gamma)) epilogue.
//   k2_stream_kernel     (wd_k2_stream.cuh) the same contraction with the mu axis streamed through a TMA-fed ring.
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
    // twist index: argmin |gamma_r|, gamma_r = D+[r] + D-[r] + lam -- restricted to the classically allowed region
    // |m| <= sqrt(j(j+1) - mu^2) + 1, where the eigenvector is O(1).  With the EXACT eigenvalue every gamma_r is pure
    // roundoff and may vanish in the exponentially small tails (|z| ~ 2^-j); twisting there makes the recurrence grow
    // by 2^j and z^2 overflows for 2j > ~1024.
    const double jj1 = 0.25 * (double)jb.twoj * (double)(jb.twoj + 2);
    const double mt = sqrt(fmax(0.0, jj1 - lam * lam)) + 1.0;
    const int i_lo = max(0, (int)ceil(0.5 * jb.twoj - mt)), i_hi = min(N - 1, (int)floor(0.5 * jb.twoj + mt));
    int r = (i_lo + i_hi) >> 1;
    double best = 1.7976931348623157e308;
    for (int i = i_lo; i <= i_hi; ++i) {
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
    int equator;                          // apply the exact zeros of :230-237 (plain kernel only)
    int handoff;                          // DMMA kernel: midpoint hand-off between paired compute warps
    int nsplit;                           // complex-D DMMA kernel: CTAs per chunk (small batches)
    long long nfull, nitems;              // real-d DMMA kernel: whole-chunk items, all items (k2_item)
    int tsplit;                           // real-d DMMA kernel: CTAs per chunk of the last partial wave
    int nostore;                          // experiment knob (WD_K2_NOSTORE=1): run the real-d DMMA kernels without their store instructions
    const int *skip_flag;                 // k2_real_kernel: device flag of k2c_pair_check_kernel; non-zero = the column kernel has this batch
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
// K4: single elements (wignerdjmn / wignerDjmn).  EIGHT lanes per tuple, four tuples per warp.
//   * Lane g of a group takes the columns kp = g, g + 8, ... of each eigenvector class of the two L2-resident table rows
//     u[m, .], u[n, .] (coalesced 64-byte pieces), so that its angle advances by a FIXED step from term to term:
//     mu -> mu + 16.  The reference takes one sincos per term (src/WignerD.jl:199); here a lane takes ONE sincos per tuple
//     (exp(i beta), or exp(i beta / 2) for half-integer j), builds the anchors of its two classes from powers of it and
//     advances with a rotation by 16 beta (4 multiply-adds) per term; beyond 16 steps (2j > 500) it re-anchors with an
//     exact sincos(fl(mu beta)).  The trig values stay within ~5e-15 of the reference's.
//   * Tried: one warp per tuple with one sin or cos per term (round 1): 6.3-6.5 ms for the 10^7 tuples of cfg-5, i.e.
//     35 % of the chip's measured FP64 sin() issue rate; one THREAD per tuple with the rows staged through shared memory:
//     7.2 ms (32 tuples of different j per warp: every warp runs to the longest of them).
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

constexpr int K4_G = 8;                     // lanes per tuple
constexpr int K4_THREADS = 256;

template <bool CPLX>
__global__ void __launch_bounds__(K4_THREADS) k4_jmn_kernel(K4Args a)
{
    const int lane = threadIdx.x & 31, g = lane & (K4_G - 1);
    const long long slots = (long long)gridDim.x * (K4_THREADS / K4_G);
    for (long long t = (long long)blockIdx.x * (K4_THREADS / K4_G) + (threadIdx.x / K4_G);; t += slots) {
        // whole warps leave together (the shuffles below are warp-wide)
        if (__all_sync(0xffffffffu, t >= a.n)) break;
        const bool live = t < a.n;
        const int tj = live ? a.twoj[t] : 0, tm = live ? a.twom[t] : 0, tn = live ? a.twon[t] : 0;
        const int kind = a.beta_kind;
        double val = 0.0;
        bool done = !live;
        // special co-latitudes (src/WignerD.jl:218-237); the poles do no index checks, like the reference
        if (live) {
            if (kind == 1) { val = (tm == tn) ? 1.0 : 0.0; done = true; }
            else if (kind == 2) { val = (tm == -tn) ? ((((tj - tn) >> 1) & 1) ? -1.0 : 1.0) : 0.0; done = true; }
            else if (kind == 3 && ((tj + tm) & 1) == 0 && ((tj + tn) & 1) == 0) {
                const bool jm_odd = ((tj + tm) >> 1) & 1, jn_odd = ((tj + tn) >> 1) & 1;
                if ((jm_odd && tn == 0) || (jn_odd && tm == 0)) { val = 0.0; done = true; }
            }
        }
        double acc0 = 0.0, acc1 = 0.0;
        bool flip1 = false;
        int sdi = 0;
        if (!done) {
            const bool ok = tj >= 0 && tj <= a.max_twoj && ((tj + tm) & 1) == 0 && ((tj + tn) & 1) == 0 &&
                            tm >= -tj && tm <= tj && tn >= -tj && tn <= tj;
            const PlanDev *pp = a.plans + (ok ? tj : 0);
            const double *uq = ok ? pp->uq : nullptr;
            if (uq == nullptr) { if (g == 0) *a.err = 1; done = true; }
            else {
                const int N = tj + 1, Q = (N + 1) >> 1, i0 = N - Q, ldu = pp->ldu, K0p = pp->K0p, Kp = pp->Kp;
                const int i = (tm + tj) >> 1, ip = (tn + tj) >> 1;
                const int qm = (i >= i0) ? i - i0 : (N - 1 - i) - i0;
                const int qn = (ip >= i0) ? ip - i0 : (N - 1 - ip) - i0;
                flip1 = ((i < i0) != (ip < i0));             // one mirrored index: class 1 changes sign
                const bool odd = (i - ip) & 1;
                sdi = i - ip;
                const double *um = uq + (size_t)qm * ldu, *un = uq + (size_t)qn * ldu;
                const int c0 = (Q - 1) & 1, par = tj & 1;
                const double beta = (kind == 3) ? 1.5707963267948966 : a.beta[t];
                // ONE sincos per lane and tuple: z = exp(i hb), hb = beta (integer j) or beta / 2 (half-integer j), and its powers
                // z^2 .. z^32 by squaring.  Every mu of the first 16 steps of a lane is a small multiple of hb (exponent <= 31), so the
                // anchors of both classes are products of those powers (<= 9 complex multiplications, ~1e-15), and the rotation
                // by 16 beta of a lane step is z^16 (z^32).  Later blocks (2j > 500) anchor with an exact sincos(fl(mu beta)) like :199.
                double pr[6], pi[6];
                sincos(par ? 0.5 * beta : beta, &pi[0], &pr[0]);
#pragma unroll
                for (int b = 1; b < 6; ++b) { pr[b] = pr[b - 1] * pr[b - 1] - pi[b - 1] * pi[b - 1]; pi[b] = 2.0 * pr[b - 1] * pi[b - 1]; }
                const double c2 = par ? pr[5] : pr[4], s2 = par ? pi[5] : pi[4];      // rotation of one lane step: mu -> mu + 16
#pragma unroll
                for (int cls = 0; cls < 2; ++cls) {
                    const int kbeg = cls ? K0p : 0, kend = cls ? Kp : K0p;     // (pad columns hold zeros)
                    const int kappa0 = (cls ? (c0 ^ 1) : c0) + 2 * g;          // mu of column kbeg + g: kappa0 (+ 1/2)
                    double acc = 0.0;
                    // blocks of 16 steps: anchor, then rotations; the loads of 4 steps are in flight together
#pragma unroll 1
                    for (int kp0 = kbeg + g, it0 = 0; kp0 < kend; kp0 += 16 * K4_G, it0 += 16) {
                        double sn, cs;
                        if (it0 == 0) {
                            const int e = par ? 2 * kappa0 + 1 : kappa0;       // mu = e * hb, e <= 31
                            cs = 1.0; sn = 0.0;
#pragma unroll
                            for (int b = 0; b < 5; ++b) {
                                const double mr = ((e >> b) & 1) ? pr[b] : 1.0, mi = ((e >> b) & 1) ? pi[b] : 0.0;
                                const double t = cs * mr - sn * mi;
                                sn = fma(sn, mr, cs * mi);
                                cs = t;
                            }
                        } else {
                            sincos(0.5 * (double)(2 * (kappa0 + 2 * K4_G * it0) + par) * beta, &sn, &cs);
                        }
                        const int nst = min(16, (kend - kp0 + K4_G - 1) / K4_G);
#pragma unroll 4
                        for (int i = 0; i < nst; ++i) {
                            const int kp = kp0 + i * K4_G;
                            const double gg = um[kp] * un[kp];
                            const double w = ((kappa0 + 2 * K4_G * (it0 + i)) | par) ? 2.0 : 1.0;      // mu = 0 has weight 1
                            acc = fma(w * (odd ? sn : cs), gg, acc);
                            const double cn = cs * c2 - sn * s2;
                            sn = fma(sn, c2, cs * s2);
                            cs = cn;
                        }
                    }
                    if (cls) acc1 = acc; else acc0 = acc;
                }
            }
        }
#pragma unroll
        for (int o = K4_G / 2; o > 0; o >>= 1) {
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
            acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
        }
        if (live && g == 0) {
            if (!done) val = flip(flip1 ? (acc0 - acc1) : (acc0 + acc1), sigma_bit(sdi));
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

}  // namespace wd
