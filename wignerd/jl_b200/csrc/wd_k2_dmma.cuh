// wd_k2_dmma.cuh -- the hot kernel: batched d^j(beta) / D^j(alpha,beta,gamma) contraction on the FP64
// tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4, the only FP64 tensor path on sm_100a).
//
//   CTA = 8 warps, persistent over chunks of 16 angles.  Shared memory: UQ (all of it, zero-padded to whole
//   16-row groups), the mu / w vectors, the chunk's trig tables Tc/Ts[16][ldt] (w cos(mu beta),
//   w sin(mu beta)), optional phase tables, and two 8 x 64 staging tiles per warp.
//   Warp unit = one quadrant column qn  x  up to GMAX groups of 16 quadrant rows  x  16 angles.
//   MMA roles:  A[8 angles][4 kp] = T(beta, kp) * UQ[qn][kp],  B[4 kp][8 rows] = UQ[row][kp],
//   C[8 angles][8 rows].  A 16-row group is two tiles -- its even rows and its odd rows -- so that one tile
//   needs one trig function; a lane then owns 4 consecutive rows of the group.  Angle tile ai holds the
//   angles of parity ai of the chunk (angle a = 2 r + ai in MMA row r), which makes the 16-byte alignment of
//   every output run of a tile the same.
//   Epilogue: P0 +- P1 butterfly and sign sigma in registers, one 8 x span tile staged in shared memory per
//   (mirror target, angle tile), then
//     real d   : the 8 runs of the tile leave through the TMA engine -- cp.async.bulk shared -> global, one bulk
//                copy of up to 512 B per run, issued by lanes 0..7; the at most two 8-byte edge elements of a
//                run that bulk copies cannot cover (16-byte alignment) are stored by lanes 8..23;
//     complex D: LDS + fused exp(-i(m alpha + n gamma)) + 16-byte STG (512 contiguous bytes per instruction).
//
//   Measured history of this kernel (cfg-2, j = 100, 10^5 angles, B200; profiles/): fully unrolled epilogue
//   -> instruction-fetch bound, 33 ms; runtime epilogue loops -> 15.4 ms; single butterfly + lean LDS/STG
//   store loop -> 12.2 ms with the LSU/shared-memory pipe as co-bottleneck of the DMMA pipe (the staged
//   read-back + STG.64 stream costs as many LSU wavefronts as the whole main loop); bulk-copy stores remove
//   that read-back from the LSU.
#pragma once
#include "wd_kernels.cuh"

namespace wd {

constexpr int K2_WARPS = 8;
constexpr int K2_THREADS = K2_WARPS * 32;
constexpr int K2_ANG = 16;                 // angles per chunk (2 MMA row tiles)
constexpr int K2_STG_LD = 66;              // staging row stride (doubles): == 2 mod 16 -> conflict-free 16-byte stores
constexpr int K2_STG_ROWS = 16;            // both angle tiles of a chunk

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d[0]), "+d"(d[1])
        : "d"(a), "d"(b));
}

// ---- TMA engine, non-tensor form: one bulk asynchronous copy global -> shared (SASS UBLKCP) brings the whole
// eigenvector table UQ of the j being computed into shared memory, completion signalled on an mbarrier.
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
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
    s.total = o;                                        // (+ 16 bytes: the unit counter)
    return s;
}

struct K2Ctx {                  // per-thread constants of the DMMA kernel
    const double *sUQ, *sTc, *sTs, *sPh;
    double *stg;
    int ldu, ldt, N, Q, i0, zskip, Kp, K0p, twoj;
    int lane, r, c;
    long long b0;
};

// row of the trig tables that holds angle a of the chunk: even angles first (tile 0), odd angles second (tile 1)
__device__ __forceinline__ int k2_trig_row(int a) { return ((a & 1) << 3) + (a >> 1); }

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

// Epilogue of one mirror target: for each angle tile, stage the 8 x span tile (signs applied) and store it.
//   DESC = false: rows ascend with the quadrant row (targets (i,i') and (i,i'r));
//   DESC = true : rows are the mirrored ones, written in ascending memory order ((ir,i'r) and (ir,i')).
// val[ai][g][eo][h]: the butterfly-combined accumulators of this target.
template <bool CPLX, bool DESC, int GMAX>
__device__ __forceinline__ void k2_store_target(const K2Args &a, K2Ctx &cx, const double (&val)[2][GMAX][2][2],
                                                const int col, const int qbase, const int gcnt)
{
    const int N = cx.N, Q = cx.Q, i0 = cx.i0, zskip = cx.zskip, lane = cx.lane, r = cx.r, c = cx.c;
    const int span = 16 * gcnt;
    // staging position s <-> global row: ascending i0+qbase+s ; descending (N-1-i0-qbase-(span-1)) + s
    const int row0 = DESC ? (N - 1 - i0 - qbase - (span - 1)) : (i0 + qbase);
    // valid staging positions [s_lo, s_hi): quadrant row < Q, and (mirrored rows) quadrant row >= zskip
    const int s_lo = DESC ? max(0, qbase + span - Q) : 0;
    const int s_hi = DESC ? min(span, qbase + span - zskip) : min(span, Q - qbase);
    if (s_hi <= s_lo) return;                                         // warp-uniform
    // sigma(row - col) of the lane's element t: row - col = +-(16g + 4c + t) + const, only mod 4 matters
    const int dib = DESC ? (N - 1 - i0 - qbase - col) : (i0 + qbase - col);
    unsigned sb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) sb[t] = sigma_bit(DESC ? dib - t : dib + t);
    const size_t NN = (size_t)N * N;
    // Stage the 16 x span tile (row ai*8 + r <-> angle 2r + ai of the chunk), then store whole column segments:
    // 256 contiguous bytes per instruction for real d, 512 for complex D (fused phase).
    double *buf = cx.stg;
    __syncwarp();
#pragma unroll
    for (int ai = 0; ai < 2; ++ai) {
#pragma unroll
        for (int g = 0; g < GMAX; ++g) {
            if (g < gcnt) {
                // the lane owns rows 16g+4c .. 16g+4c+3 of the unit: (even tile, odd tile) interleaved
                const double v0 = flip(val[ai][g][0][0], sb[0]), v1 = flip(val[ai][g][1][0], sb[1]);
                const double v2 = flip(val[ai][g][0][1], sb[2]), v3 = flip(val[ai][g][1][1], sb[3]);
                double *p = buf + (ai * 8 + r) * K2_STG_LD + (DESC ? (span - 4) - 16 * g - 4 * c : 16 * g + 4 * c);
                reinterpret_cast<double2 *>(p)[0] = DESC ? make_double2(v3, v2) : make_double2(v0, v1);
                reinterpret_cast<double2 *>(p)[1] = DESC ? make_double2(v1, v0) : make_double2(v2, v3);
            }
        }
    }
    __syncwarp();
    const size_t e00 = (size_t)cx.b0 * NN + (size_t)col * N + row0;   // angle 0 of the chunk, staging position 0
#pragma unroll 1
    for (int hh = 0; hh * 32 < span; ++hh) {
        const int s = hh * 32 + lane;
        if (s < s_lo || s >= s_hi) continue;
        const double *src = buf + s;
        if (!CPLX) {
#pragma unroll
            for (int ai = 0; ai < 2; ++ai) {
                const long long bfirst = cx.b0 + ai;
                if (bfirst >= a.nb) break;                            // warp-uniform
                const int nrows = (int)min(8LL, (a.nb - bfirst + 1) / 2);
                const double *sp = src + ai * 8 * K2_STG_LD;
                double *dst = a.out + e00 + (size_t)ai * NN + s;
#pragma unroll 4
                for (int ar = 0; ar < nrows; ++ar) {                  // angle 2 ar + ai of the chunk
                    *dst = *sp;
                    sp += K2_STG_LD;
                    dst += 2 * NN;
                }
            }
        } else {
            // exp(-i(m alpha + n gamma)) = (ca - i sa)(cg - i sg); m -> -m flips sa, n -> -n flips sg
            const int row = row0 + s;
            const bool mneg = row < i0, nneg = col < i0;
            const int qr = mneg ? (N - 1 - row - i0) : (row - i0);
            const int qc = nneg ? (N - 1 - col - i0) : (col - i0);
#pragma unroll
            for (int ai = 0; ai < 2; ++ai) {
                const long long bfirst = cx.b0 + ai;
                if (bfirst >= a.nb) break;                            // warp-uniform
                const int nrows = (int)min(8LL, (a.nb - bfirst + 1) / 2);
                const double *pa = cx.sPh + ai * Q + qr, *pg = cx.sPh + (2 * K2_ANG + ai) * Q + qc;
                const double *sp = src + ai * 8 * K2_STG_LD;
                double2 *dst = reinterpret_cast<double2 *>(a.out) + e00 + (size_t)ai * NN + s;
#pragma unroll 2
                for (int ar = 0; ar < nrows; ++ar) {                  // angle 2 ar + ai of the chunk
                    const double v = sp[ar * K2_STG_LD];
                    const double ca = pa[(2 * ar) * Q], cg = pg[(2 * ar) * Q];
                    double sa = pa[(K2_ANG + 2 * ar) * Q], sg = pg[(K2_ANG + 2 * ar) * Q];
                    if (mneg) sa = -sa;
                    if (nneg) sg = -sg;
                    const double cr = ca * cg - sa * sg;
                    const double ci = -(sa * cg + ca * sg);
                    dst[(size_t)(2 * ar) * NN] = make_double2(v * cr, v * ci);
                }
            }
        }
    }
}

// One warp unit: quadrant column qn, gcnt groups of 16 quadrant rows starting at group g0, 16 angles.
template <bool HALF, bool CPLX>
__device__ __forceinline__ void k2_unit(const K2Args &a, K2Ctx &cx, const int qn, const int g0, const int gcnt)
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
    int *sCounter = reinterpret_cast<int *>(sm + L.total);     // one int after the staging tiles
    const int tid = threadIdx.x, warp = tid >> 5;
    K2Ctx cx;
    cx.sUQ = sUQ; cx.sTc = sTc; cx.sTs = sTs; cx.sPh = sPh;
    cx.stg = sm + L.stg + warp * (K2_STG_ROWS * K2_STG_LD);
    cx.ldu = ldu; cx.ldt = ldt; cx.N = p.N; cx.Q = Q; cx.i0 = p.N - Q; cx.zskip = p.zskip;
    cx.Kp = Kp; cx.K0p = p.K0p; cx.twoj = p.twoj;
    cx.lane = tid & 31; cx.r = cx.lane >> 2; cx.c = cx.lane & 3;

    // one-time: UQ -> shared through the TMA engine (one bulk copy, mbarrier-signalled); zero pad rows, mu, w by hand
    {
        void *mbar = reinterpret_cast<char *>(sm + L.total) + 8;   // 8 bytes after the unit counter
        const unsigned uq_bytes = (unsigned)(Q * ldu) * 8u;        // a multiple of 16: ldu is even
        if (tid == 0) {
            mbar_init(mbar, 1);
            mbar_expect_tx(mbar, uq_bytes);
            bulk_load(sUQ, p.uq, uq_bytes, mbar);
        }
        for (int i = Q * ldu + tid; i < L.qpad * ldu; i += K2_THREADS) sUQ[i] = 0.0;
        for (int i = tid; i < Kp; i += K2_THREADS) { sMu[i] = p.mu[i]; sW[i] = p.w[i]; }
        __syncthreads();                                           // the mbarrier is initialised for everybody
        mbar_wait(mbar, 0);
    }
    const int ngroups = L.qpad >> 4;
    const int upc = (ngroups + GMAX - 1) / GMAX;          // parts per column
    const int gbase = ngroups / upc, grem = ngroups - gbase * upc;   // the first grem parts have gbase+1 groups
    const int nunits = Q * upc;
    const long long nchunks = (a.nb + K2_ANG - 1) / K2_ANG;

    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long b0 = ch * K2_ANG;
        cx.b0 = b0;
        __syncthreads();                                   // previous chunk's readers are done (and UQ is in)
        if (tid == 0) *sCounter = 0;
#pragma unroll 1
        for (int idx = tid; idx < K2_ANG * Kp; idx += K2_THREADS) {
            const int ang = idx / Kp, kp = idx - ang * Kp;
            long long b = b0 + ang; if (b >= a.nb) b = a.nb - 1;
            double s, cs;
            sincos(sMu[kp] * a.beta[b], &s, &cs);         // fl(mu*beta) then sincos, like :199
            const double w = sW[kp];
            const int row = k2_trig_row(ang);
            sTc[row * ldt + kp] = w * cs;
            sTs[row * ldt + kp] = w * s;
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

        // dynamic unit scheduler (one shared counter per chunk): unit v -> part v / Q of column v % Q, so the
        // larger parts of all columns are handed out before the smaller ones
#pragma unroll 1
        for (;;) {
            int v = 0;
            if (cx.lane == 0) v = atomicAdd(sCounter, 1);
            v = __shfl_sync(0xffffffffu, v, 0);
            if (v >= nunits) break;
            const int part = v / Q, qn = v - part * Q;
            const int g0 = part * gbase + min(part, grem);
            const int gcnt = gbase + (part < grem ? 1 : 0);
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
