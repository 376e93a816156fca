// wd_k2_dmma.cuh -- the hot kernel: batched d^j(beta) / D^j(alpha,beta,gamma) contraction on the FP64
// tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4, the only FP64 tensor path on sm_100a).
//
//   CTA = 8 warps, persistent over chunks of 16 angles.  Shared memory: UQ (all of it, zero-padded to whole
//   16-row groups), the mu / w vectors, the chunk's trig tables Tc/Ts[16][ldt] (w cos(mu beta),
//   w sin(mu beta)), optional phase tables, and one 16 x 64 staging tile per warp.
//   Warp unit = one quadrant column qn  x  up to GMAX groups of 16 quadrant rows  x  16 angles.
//   MMA roles:  A[8 angles][4 kp] = T(beta, kp) * UQ[qn][kp],  B[4 kp][8 rows] = UQ[row][kp],
//   C[8 angles][8 rows].  A 16-row group is two tiles -- its even rows and its odd rows -- so that one tile
//   needs one trig function; a lane then owns 4 consecutive rows of the group.
//   Epilogue: P0 +- P1 butterfly, sign sigma, staged through shared memory so that every global store
//   instruction writes 256 contiguous bytes of one output column (512 for complex D).
//
//   Code-size note (ncu, round 1): a fully unrolled epilogue made this kernel 340 KB of SASS and
//   instruction-fetch bound ("no_instruction" 60 % of stall samples); the epilogue is therefore a runtime
//   loop over the four mirror targets and the main loop is the only part specialised on the group count.
#pragma once
#include "wd_kernels.cuh"

namespace wd {

constexpr int K2_WARPS = 8;
constexpr int K2_THREADS = K2_WARPS * 32;
constexpr int K2_ANG = 16;                 // angles per chunk (2 MMA row tiles)
constexpr int K2_STG_LD = 66;              // staging row stride (doubles): == 2 mod 4, 16B-aligned rows
constexpr int K2_STG_ROWS = 16;

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

// Main loop of one warp unit, specialised on the number GC of 16-row groups actually present.
// acc[class][angle tile][group][even/odd rows][main/alt][frag]; groups >= GC stay untouched (zero).
template <bool HALF, int GMAX, int GC>
__device__ __forceinline__ void k2_mainloop(double (&acc)[2][2][GMAX][2][HALF ? 2 : 1][2], const K2Ctx &cx,
                                            const int qn, const int g0)
{
    const int ldu = cx.ldu, ldt = cx.ldt, r = cx.r, c = cx.c;
    const int pe = qn & 1;
    // even-row tiles pair with cos when qn is even, with sin when qn is odd
    const double *tA = (pe ? cx.sTs : cx.sTc) + r * ldt + c;     // trig for the even-row tiles
    const double *tB = (pe ? cx.sTc : cx.sTs) + r * ldt + c;     // trig for the odd-row tiles
    const double *un = cx.sUQ + qn * ldu + c;
    const double *ub = cx.sUQ + (16 * g0 + 2 * r) * ldu + c;     // even rows; odd rows at +ldu; groups at +16*ldu
    const int gs = 16 * ldu, t8 = 8 * ldt;
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
#pragma unroll 2
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
}

// Epilogue of one mirror target: stage the warp's 16 x span tile through shared memory, then store whole
// columns segments (256 contiguous bytes per instruction; 512 for complex D).
//   DESC = false: rows ascend with the quadrant row (targets (i,i') and (i,i'r));
//   DESC = true : rows are the mirrored ones, written in ascending memory order ((ir,i'r) and (ir,i')).
// val[ai][g][eo][h]: the butterfly-combined accumulators of this target.
template <bool CPLX, bool DESC, int GMAX>
__device__ __forceinline__ void k2_store_target(const K2Args &a, const K2Ctx &cx, const double (&val)[2][GMAX][2][2],
                                                const int col, const int qbase, const int gcnt)
{
    const int N = cx.N, Q = cx.Q, i0 = cx.i0, zskip = cx.zskip, lane = cx.lane, r = cx.r, c = cx.c;
    double *stg = cx.stg;
    const int span = 16 * gcnt;
    __syncwarp();
#pragma unroll
    for (int ai = 0; ai < 2; ++ai) {
#pragma unroll
        for (int g = 0; g < GMAX; ++g) {
            if (g < gcnt) {
                // the lane owns rows 16g+4c .. 16g+4c+3 of the unit: (even tile, odd tile) interleaved
                const double v0 = val[ai][g][0][0], v1 = val[ai][g][1][0], v2 = val[ai][g][0][1], v3 = val[ai][g][1][1];
                if (!DESC) {
                    double2 *dst = reinterpret_cast<double2 *>(stg + (ai * 8 + r) * K2_STG_LD + 16 * g + 4 * c);
                    dst[0] = make_double2(v0, v1);
                    dst[1] = make_double2(v2, v3);
                } else {
                    double2 *dst = reinterpret_cast<double2 *>(stg + (ai * 8 + r) * K2_STG_LD + (span - 4) - 16 * g - 4 * c);
                    dst[0] = make_double2(v3, v2);
                    dst[1] = make_double2(v1, v0);
                }
            }
        }
    }
    __syncwarp();
    // staging position s <-> global row: ascending i0+qbase+s ; descending (N-1-i0-qbase-(span-1)) + s
    const int row0 = DESC ? (N - 1 - i0 - qbase - (span - 1)) : (i0 + qbase);
    const int nrows = (int)min((long long)K2_ANG, a.nb - cx.b0);
    const size_t mstride = (size_t)N * N * (CPLX ? 2 : 1);
#pragma unroll 1
    for (int hh = 0; hh * 32 < span; ++hh) {
        const int s = hh * 32 + lane;
        const int qm = DESC ? (qbase + span - 1 - s) : (qbase + s);
        if (s >= span || qm >= Q || (DESC && qm < zskip)) continue;
        const int row = row0 + s;
        unsigned sgn = sigma_bit(row - col);                      // sigma(i - i')
        unsigned keep = 0xffffffffu;
        if (a.equator && zskip) {
            const int j = cx.twoj >> 1;
            if ((col == j && (row & 1)) || (row == j && (col & 1))) { keep = 0u; sgn = 0u; }   // exact zeros of :230-237
        }
        const double *src = stg + s;
        double *dst = a.out + ((size_t)cx.b0 * N * N + (size_t)col * N + row) * (CPLX ? 2 : 1);
        if (!CPLX) {
#pragma unroll 4
            for (int ar = 0; ar < nrows; ++ar) {
                const double v = src[ar * K2_STG_LD];
                dst[ar * mstride] = __hiloint2double((__double2hiint(v) ^ (int)sgn) & (int)keep, __double2loint(v) & (int)keep);
            }
        } else {
            // exp(-i(m alpha + n gamma)) = (ca - i sa)(cg - i sg); m -> -m flips sa, n -> -n flips sg
            const bool mneg = row < i0, nneg = col < i0;
            const int qr = mneg ? (N - 1 - row - i0) : (row - i0);
            const int qc = nneg ? (N - 1 - col - i0) : (col - i0);
            const double *pa = cx.sPh + qr, *pg = cx.sPh + 2 * K2_ANG * Q + qc;
#pragma unroll 2
            for (int ar = 0; ar < nrows; ++ar) {
                double v = src[ar * K2_STG_LD];
                v = __hiloint2double((__double2hiint(v) ^ (int)sgn) & (int)keep, __double2loint(v) & (int)keep);
                const double ca = pa[ar * Q], cg = pg[ar * Q];
                double sa = pa[(K2_ANG + ar) * Q], sg = pg[(K2_ANG + ar) * Q];
                if (mneg) sa = -sa;
                if (nneg) sg = -sg;
                const double cr = ca * cg - sa * sg;
                const double ci = -(sa * cg + ca * sg);
                *reinterpret_cast<double2 *>(dst + ar * mstride) = make_double2(v * cr, v * ci);
            }
        }
    }
}

// One warp unit: quadrant column qn, gcnt groups of 16 quadrant rows starting at group g0, 16 angles.
template <bool HALF, bool CPLX>
__device__ __forceinline__ void k2_unit(const K2Args &a, const K2Ctx &cx, const int qn, const int g0, const int gcnt)
{
    constexpr int GMAX = HALF ? 2 : 4;
    constexpr int NALT = HALF ? 2 : 1;
    double acc[2][2][GMAX][2][NALT][2];
#pragma unroll
    for (int i = 0; i < 2 * 2 * GMAX * 2 * NALT * 2; ++i) (&acc[0][0][0][0][0][0])[i] = 0.0;

    if (HALF) {
        if (gcnt == 2) k2_mainloop<HALF, GMAX, (GMAX >= 2 ? 2 : 1)>(acc, cx, qn, g0);
        else k2_mainloop<HALF, GMAX, 1>(acc, cx, qn, g0);
    } else {
        switch (gcnt) {
        case 4: k2_mainloop<HALF, GMAX, (GMAX >= 4 ? 4 : 1)>(acc, cx, qn, g0); break;
        case 3: k2_mainloop<HALF, GMAX, (GMAX >= 4 ? 3 : 1)>(acc, cx, qn, g0); break;
        case 2: k2_mainloop<HALF, GMAX, (GMAX >= 2 ? 2 : 1)>(acc, cx, qn, g0); break;
        default: k2_mainloop<HALF, GMAX, 1>(acc, cx, qn, g0); break;
        }
    }

    // ---- epilogue ---------------------------------------------------------------------------------------
    // butterfly, once: sum = P0 + P1 (targets (i,i'), (ir,i'r)), dif = P0 - P1 (targets (ir,i'), (i,i'r));
    // for half-integer j the mirrored pair uses the accumulators of the other trig function (alt).
    double vs[2][GMAX][2][2], vd[2][GMAX][2][2];
#pragma unroll
    for (int ai = 0; ai < 2; ++ai)
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
#pragma unroll
            for (int eo = 0; eo < 2; ++eo)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    vs[ai][g][eo][h] = acc[0][ai][g][eo][0][h] + acc[1][ai][g][eo][0][h];
                    vd[ai][g][eo][h] = acc[0][ai][g][eo][NALT - 1][h] - acc[1][ai][g][eo][NALT - 1][h];
                }
    const int ip = cx.i0 + qn, ipr = cx.N - 1 - ip;
    const int qbase = 16 * g0;                     // first quadrant row of the unit
    k2_store_target<CPLX, false, GMAX>(a, cx, vs, ip, qbase, gcnt);                 // ( i, i')
    k2_store_target<CPLX, true, GMAX>(a, cx, vd, ip, qbase, gcnt);                  // (ir, i')
    if (qn >= cx.zskip) {                          // n = 0 mirrors onto itself (warp-uniform)
        k2_store_target<CPLX, true, GMAX>(a, cx, vs, ipr, qbase, gcnt);             // (ir, i'r)
        k2_store_target<CPLX, false, GMAX>(a, cx, vd, ipr, qbase, gcnt);            // ( i, i'r)
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
#pragma unroll 1
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
#pragma unroll 1
            for (int idx = tid; idx < 2 * K2_ANG * Q; idx += K2_THREADS) {
                const int which = idx / (K2_ANG * Q), rem = idx - which * (K2_ANG * Q);
                const int ang = rem / Q, q = rem - ang * Q;
                long long b = b0 + ang; if (b >= a.nb) b = a.nb - 1;
                const double m = 0.5 * (double)(2 * q + (p.twoj & 1));
                double s, cs;
                sincos(m * (which ? a.gamma[b] : a.alpha[b]), &s, &cs);
                sPh[((2 * which + 0) * K2_ANG + ang) * Q + q] = cs;
                sPh[((2 * which + 1) * K2_ANG + ang) * Q + q] = s;
            }
        }
        __syncthreads();

#pragma unroll 1
        for (int u = warp; u < nunits; u += K2_WARPS) {
            const int qn = u / upc;
            const int g0 = (u - qn * upc) * GMAX;
            const int gcnt = min(GMAX, ngroups - g0);
            k2_unit<HALF, CPLX>(a, cx, qn, g0, gcnt);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FP64 peak micro-benchmarks (roofline calibration)
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
