// wd_k2_dmma.cuh -- the hot kernels: batched d^j(beta) / D^j(alpha,beta,gamma) contraction on the FP64
// tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4, the only FP64 tensor path on sm_100a), table of the j being
// computed resident in shared memory (2j <= 223; wd_k2_stream.cuh streams the mu axis beyond).
//
//   k2_real_kernel<HALF>  real d.  16 warps, warp-specialised: 8 COMPUTE warps (200 registers after setmaxnreg.inc)
//                         and 8 STORE warps (56 after setmaxnreg.dec).  Persistent over chunks of 16 angles; the next
//                         chunk's trig tables are made while the current one is computed; no CTA barrier.
//   k2_cplx_kernel<HALF>  complex D.  8 warps that do both jobs (the drain multiplies by the per-chunk phase tables).
//
//   Shared memory: UQ (all of it, one TMA bulk copy, zero-padded to whole 16-row groups), the mu / w vectors, the
//   trig tables Tc/Ts[16][ldt] (w cos(mu beta), w sin(mu beta); two buffers in the real kernel), the phase tables
//   (complex), and per compute warp a ring of two 8 x 64 staging slots + descriptors + full/empty mbarriers.
//   Warp unit = one quadrant column qn  x  up to GMAX groups of 16 quadrant rows  x  16 angles.
//   MMA roles:  A[8 angles][4 kp] = T(beta, kp) * UQ[qn][kp],  B[4 kp][8 rows] = UQ[row][kp],
//   C[8 angles][8 rows].  A 16-row group is two tiles -- its even rows and its odd rows -- so that one tile
//   needs one trig function; a lane then owns 4 consecutive rows of the group.  Angle tile ai holds the
//   angles of parity ai of the chunk (angle a = 2 r + ai in MMA row r).
//   Epilogue: butterfly S = P0 + P1, D = P0 - P1.  S belongs to the mirror pair (i,i'), (ir,i'r), D to (ir,i'),
//   (i,i'r): each is staged ONCE per angle tile (conflict-free 16-byte STS) and every staged value is written to both
//   targets of its pair -- an ascending run in one output column, a descending run in the mirror column -- by the
//   store warp (real d: sign sigma in place, 256 contiguous bytes per store instruction) or by the warp itself
//   (complex D: conjugate phases, 512 bytes per instruction).  Partial-sector writes are what the L2 cannot absorb:
//   direct register stores of the same data ran the first version of this kernel at 17-19 ms.
//
//   Why warp specialisation (ncu + experiment builds, profiles/): with 8 symmetric warps the drain loop --
//   ~1100 dependent integer/LDS/STG instructions per unit -- kept every warp out of its main loop for ~45 % of the
//   time (12.1 ms at cfg-2); without the drain the same kernel ran in 8.9 ms.  DESIGN.md section 4 has the history.
#pragma once
#include "wd_kernels.cuh"

namespace wd {

constexpr int K2_CWARPS = 8;               // compute warps
constexpr int K2_SWARPS = 8;               // store warps, one per compute warp (real-d kernels only)
constexpr int K2_ANG = 16;                 // angles per chunk (2 MMA row tiles)
constexpr int K2_STG_LD = 66;              // staging row stride (doubles): == 2 mod 16 -> conflict-free 16-byte stores
constexpr int K2_HT_ROWS = 8;               // angles per staged tile (one MMA row tile = "half tile" of a chunk)
constexpr int K2_HT_SIZE = K2_HT_ROWS * K2_STG_LD;   // doubles per staged tile
constexpr int K2_REGS_COMPUTE = 200;
constexpr int K2_REGS_STORE = 56;             // 8*32*200 + 8*32*56 = 65536 = the whole register file
constexpr int K2_THREADS_REAL = (K2_CWARPS + K2_SWARPS) * 32;      // real-d kernels: compute + store warps
constexpr int K2_THREADS_CPLX = K2_CWARPS * 32;                    // complex-D kernel: every warp does both jobs

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d[0]), "+d"(d[1])
        : "d"(a), "d"(b));
}

// ---- mbarriers, and the TMA engine in its non-tensor form: one bulk asynchronous copy global -> shared (SASS
// UBLKCP) brings the whole eigenvector table UQ of the j being computed into shared memory.
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_test(void *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    while (!mbar_test(bar, parity)) { }
}
// named barriers (bar.* is warp-aligned: the warp must be converged, hence the __syncwarp)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads)
{
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct K2Smem {                 // complex-D kernel; offsets in doubles, computed identically on host and device
    int uq, mu, w, tc, ts, ph, stg, bars, total;
    int ldt, qpad;
};
__host__ __device__ inline K2Smem k2_smem_layout(int Q, int Kp, int ldu)
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
    s.ph = o;  o += 4 * K2_ANG * Q;                     // cos(m a), sin(m a), cos(n g), sin(n g)  [16][Q]
    o = (o + 1) & ~1;
    s.stg = o;  o += K2_CWARPS * K2_HT_ROWS * K2_STG_LD;   // one angle tile (8 angles) per compute warp
    s.bars = o; o += 2;                                 // UQ-load barrier, unit counter
    s.total = o;
    return s;
}

struct K2Ctx {                  // per-thread constants of a compute warp
    const double *sUQ, *sTc, *sTs, *sPh;
    double *stg;                // complex-D kernel: the warp's staging tile
    int ldu, ldt, N, Q, i0, zskip, Kp, K0p, twoj;
    int lane, r, c;
    int bar_me, bar_other;      // named barriers of the hand-off between the two warps of a sub-partition (0 = off)
    int krel;                   // release point: after this many pairs of k-steps of the first class
    long long b0;
};

// row of the trig tables that holds angle a of the chunk: even angles first (tile 0), odd angles second (tile 1)
__device__ __forceinline__ int k2_trig_row(int a) { return ((a & 1) << 3) + (a >> 1); }

// trig tables of one chunk: Tc/Ts[row(angle)][kp] = w cos|sin(mu beta), fl(mu*beta) then sincos like src/WignerD.jl:199
__device__ __forceinline__ void k2_trig_fill(const K2Args &a, const double *sMu, const double *sW, double *tc, const int ldt,
                                             const int Kp, const long long b0, const int t0, const int nt)
{
    double *ts = tc + K2_ANG * ldt;
#pragma unroll 1
    for (int idx = t0; idx < K2_ANG * Kp; idx += nt) {
        const int ang = idx / Kp, kp = idx - ang * Kp;
        long long b = b0 + ang; if (b >= a.nb) b = a.nb - 1;
        double sn, cs;
        sincos(sMu[kp] * a.beta[b], &sn, &cs);
        const double w = sW[kp];
        const int row = k2_trig_row(ang);
        tc[row * ldt + kp] = w * cs;
        ts[row * ldt + kp] = w * sn;
    }
}

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
        // Software pipeline with two explicit fragment sets (no register moves): while the 16 DMMAs of one k-step
        // issue from set X, the loads of the next step land in set Y.  A load past the last step reads (and
        // discards) in-bounds shared memory.
        struct Frag { double un, ta[2], tb[2], be[GC], bo[GC]; };
        auto load = [&](Frag &f, const int k) {
            f.un = un[k];
#pragma unroll
            for (int ai = 0; ai < 2; ++ai) { f.ta[ai] = tA[ai * t8 + k]; f.tb[ai] = tB[ai * t8 + k]; }
#pragma unroll
            for (int g = 0; g < GC; ++g) { f.be[g] = ub[g * gs + k]; f.bo[g] = ub[g * gs + ldu + k]; }
        };
        auto step = [&](Frag &f) {
#pragma unroll
            for (int ai = 0; ai < 2; ++ai) { f.ta[ai] *= f.un; f.tb[ai] *= f.un; }      // A = T(beta, kp) * UQ[qn][kp]
#pragma unroll
            for (int g = 0; g < GC; ++g) {
#pragma unroll
                for (int ai = 0; ai < 2; ++ai) {
                    dmma(acc[cls][ai][g][0][0], f.ta[ai], f.be[g]);
                    dmma(acc[cls][ai][g][1][0], f.tb[ai], f.bo[g]);
                    if (HALF) {
                        dmma(acc[cls][ai][g][0][1], f.tb[ai], f.be[g]);
                        dmma(acc[cls][ai][g][1][1], f.ta[ai], f.bo[g]);
                    }
                }
            }
        };
        Frag f0, f1;
        load(f0, kbeg);
        int k = kbeg;
        // hand-off: the partner warp is released after cx.krel pairs of k-steps of class 0 (or at its end)
        // (krel >= 100: in class 1, after krel - 100 pairs)
        const int krel = !cx.bar_other ? -1 : (cls == 0 ? (cx.krel < 100 ? kbeg + 8 * cx.krel : -1)
                                                           : (cx.krel >= 100 ? kbeg + 8 * (cx.krel - 100) : -1));
        bool released = false;
#pragma unroll 1
        for (; k + 4 < kend; k += 8) {
            if (k == krel) { named_bar_arrive(cx.bar_other, 64); released = true; }
            load(f1, k + 4);
            step(f0);
            load(f0, k + 8);
            step(f1);
        }
        if (k < kend) step(f0);
        if (cx.bar_other && !released && ((cls == 0 && cx.krel < 100) || (cls == 1 && cx.krel >= 100))) named_bar_arrive(cx.bar_other, 64);
    }
}

// ------------------------------------------------------------------------------------------------
// Complex D (k2_cplx_kernel): 8 warps that compute AND store (the drain multiplies by the per-chunk phase tables;
// a 56-register store warp is too slow for that: 7.4 vs 5.4 ms at cfg-3).
// Epilogue of one butterfly result (S = P0 + P1 or D = P0 - P1): staged once per angle tile, every staged value is
// written to its two mirror targets with the fused phase exp(-i(m alpha + n gamma)) = (ca - i sa)(cg - i sg):
//     S: (i,i')  m,n >= 0: re = x - y, im = -(z + w)        (ir,i'r) m,n <= 0: re = x - y, im = +(z + w)
//     D: (i,i'r) m >= 0 >= n: re = x + y, im = w - z         (ir,i')  m <= 0 <= n: re = x + y, im = z - w
// with x = ca cg, y = sa sg, z = sa cg, w = ca sg for m, n >= 0: the two targets of a tile are complex conjugates
// up to the sign sigma.  val[ai][g][eo][h]: the butterfly-combined accumulators.
template <int GMAX>
__device__ __forceinline__ void k2_emit_cplx(const K2Args &a, const K2Ctx &cx, const double (&val)[2][GMAX][2][2],
                                             const bool isD, const int qn, const int qbase, const int gcnt)
{
    const int N = cx.N, Q = cx.Q, i0 = cx.i0, lane = cx.lane;
    const int ip = i0 + qn, ipr = N - 1 - ip;
    const int col_a = isD ? ipr : ip, col_d = isD ? ip : ipr;       // S: (i,i') up, (ir,i'r) down;  D: (i,i'r) up, (ir,i') down
    const bool both = qn >= cx.zskip;                               // n = 0 mirrors onto itself
    const bool on_a = (!isD || both) && !a.nostore, on_d0 = (isD || both) && !a.nostore;     // (nostore: experiment knob)
    const int row_a = i0 + qbase, row_d = N - 1 - i0 - qbase;
    const int s_hi = min(16 * gcnt, Q - qbase), s_lo_d = max(0, cx.zskip - qbase);
    const int dib_a = row_a - col_a, dib_d = row_d - col_d;
    const size_t NN = (size_t)N * N;
    double2 *out = reinterpret_cast<double2 *>(a.out);
    double *buf = cx.stg;
#pragma unroll
    for (int ai = 0; ai < 2; ++ai) {
        const long long bf = cx.b0 + ai;                              // first angle of the tile
        if (bf >= a.nb) break;                                        // warp-uniform
        const int nrows = (int)min(8LL, (a.nb - bf + 1) / 2);
        __syncwarp();                                                 // the previous drain has read the buffer
        {
            double *p = buf + cx.r * K2_STG_LD + 4 * cx.c;            // the lane owns rows 16g+4c .. 16g+4c+3 of the unit
#pragma unroll
            for (int g = 0; g < GMAX; ++g) {
                if (g < gcnt) {
                    reinterpret_cast<double2 *>(p + 16 * g)[0] = make_double2(val[ai][g][0][0], val[ai][g][1][0]);
                    reinterpret_cast<double2 *>(p + 16 * g)[1] = make_double2(val[ai][g][0][1], val[ai][g][1][1]);
                }
            }
        }
        __syncwarp();
        const double *pg = cx.sPh + (2 * K2_ANG + ai) * Q + qn;       // cos(n gamma) of angle ai; sin at + K2_ANG rows
        double2 *base_a = out + (size_t)bf * NN + (size_t)col_a * N + row_a;
        double2 *base_d = out + (size_t)bf * NN + (size_t)col_d * N + row_d;
        const int pstep = 2 * Q, psin = K2_ANG * Q;                   // next angle of the tile; cos -> sin table
        const size_t dstep = 2 * NN;
#pragma unroll 1
        for (int s = lane; s < s_hi; s += 32) {
            const double *pa = cx.sPh + ai * Q + qbase + s;           // cos(m alpha) of angle ai; sin at + K2_ANG rows
            const double *pgg = pg;
            const double *sp = buf + s;
            const bool on_d = on_d0 && s >= s_lo_d;
            const unsigned sga = sigma_bit(dib_a + s), sgd = sigma_bit(dib_d - s) ^ 0x80000000u;   // conjugate: im flips
            double2 *da = base_a + s, *dd = base_d - s;
            // (explicit pointer stepping: with indexed addressing the loop spent ~20 integer instructions per angle)
#pragma unroll 4
            for (int ar = 0; ar < nrows; ++ar) {                      // angle 2 ar + ai of the chunk
                const double v = *sp;
                const double ca = pa[0], sa = pa[psin];
                const double cg = pgg[0], sg = pgg[psin];
                sp += K2_STG_LD; pa += pstep; pgg += pstep;
                const double x = ca * cg, y = sa * sg, z = sa * cg, w = ca * sg;
                const double re = isD ? x + y : x - y;
                const double im = isD ? w - z : -(z + w);
                const double vr = v * re, vi = v * im;
                if (on_a) *da = make_double2(flip(vr, sga), flip(vi, sga));
                if (on_d) *dd = make_double2(flip(vr, sgd ^ 0x80000000u), flip(vi, sgd));
                da += dstep; dd += dstep;
            }
        }
    }
}

// One warp unit: quadrant column qn, gcnt groups of 16 quadrant rows starting at group g0, 16 angles.
template <bool HALF>
__device__ __forceinline__ void k2_unit_cplx(const K2Args &a, const K2Ctx &cx, const int qn, const int g0, const int gcnt)
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
    // butterfly, once: S = P0 + P1 (targets (i,i'), (ir,i'r)), D = P0 - P1 (targets (ir,i'), (i,i'r));
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
    k2_emit_cplx<GMAX>(a, cx, vs, false, qn, 16 * g0, gcnt);
    k2_emit_cplx<GMAX>(a, cx, vd, true, qn, 16 * g0, gcnt);
}

template <bool HALF>
__global__ void __launch_bounds__(K2_THREADS_CPLX, 1) k2_cplx_kernel(K2Args a)
{
    constexpr int GMAX = HALF ? 2 : 4;
    constexpr int NC = K2_CWARPS * 32;
    extern __shared__ __align__(128) double sm[];
    const PlanDev &p = a.p;
    const int Q = p.Q, Kp = p.Kp, ldu = p.ldu;
    const K2Smem L = k2_smem_layout(Q, Kp, ldu);
    const int ldt = L.ldt;
    double *sUQ = sm + L.uq, *sMu = sm + L.mu, *sW = sm + L.w, *sTc = sm + L.tc, *sTs = sm + L.ts;
    double *sPh = sm + L.ph;
    void *uq_bar = sm + L.bars;
    int *sCounter = reinterpret_cast<int *>(sm + L.bars + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // one-time: UQ -> shared through the TMA engine (one bulk copy, mbarrier-signalled); zero pad rows, mu, w by hand
    {
        const unsigned uq_bytes = (unsigned)(Q * ldu) * 8u;        // a multiple of 16: ldu is even
        if (tid == 0) {
            mbar_init(uq_bar, 1);
            mbar_init_fence();
            mbar_expect_tx(uq_bar, uq_bytes);
            bulk_load(sUQ, p.uq, uq_bytes, uq_bar);
        }
        for (int i = Q * ldu + tid; i < L.qpad * ldu; i += NC) sUQ[i] = 0.0;
        for (int i = tid; i < Kp; i += NC) { sMu[i] = p.mu[i]; sW[i] = p.w[i]; }
        __syncthreads();                                           // the mbarrier is initialised for everybody
        mbar_wait(uq_bar, 0);
    }

    K2Ctx cx;
    cx.sUQ = sUQ; cx.sTc = sTc; cx.sTs = sTs; cx.sPh = sPh;
    cx.stg = sm + L.stg + warp * (K2_HT_ROWS * K2_STG_LD);
    cx.ldu = ldu; cx.ldt = ldt; cx.N = p.N; cx.Q = Q; cx.i0 = p.N - Q; cx.zskip = p.zskip;
    cx.Kp = Kp; cx.K0p = p.K0p; cx.twoj = p.twoj;
    cx.lane = lane; cx.r = lane >> 2; cx.c = lane & 3;
    // Anti-phase scheduling of the two warps (w, w+4) that share an SM sub-partition and hence one DMMA pipe: a warp
    // may start a main loop only after its partner has passed a release point of its own (one token, a pair of named
    // barriers), so the main loops overlap only partly and the epilogues of the pair never coincide.  (Left alone the
    // pair locks into phase: the warp that is alone in its main loop runs faster than half speed, so any offset
    // shrinks, and then both are in their epilogues while the pipe idles.)
    cx.krel = a.handoff - 1;
    cx.bar_me = a.handoff ? 2 + warp : 0;
    cx.bar_other = a.handoff ? 2 + (warp ^ 4) : 0;
    if (a.handoff && (warp & 4)) named_bar_arrive(cx.bar_other, 64);       // warps 0..3 take the first turn

    const int ngroups = L.qpad >> 4;
    const int upc = (ngroups + GMAX - 1) / GMAX;          // parts per column
    const int gbase = ngroups / upc, grem = ngroups - gbase * upc;   // the first grem parts have gbase+1 groups
    const int nunits = Q * upc;
    const long long nchunks = (a.nb + K2_ANG - 1) / K2_ANG;

    // work item = (chunk, split): small batches are split over several CTAs per chunk so that the whole GPU is used
    const int nsplit = max(1, a.nsplit);
    for (long long item = blockIdx.x; item < nchunks * nsplit; item += gridDim.x) {
        const long long ch = item / nsplit;
        const int split = (int)(item - ch * nsplit);
        const long long b0 = ch * K2_ANG;
        cx.b0 = b0;
        named_bar_sync(1, NC);                             // the previous chunk's table readers are done
        const int u_begin = (int)((long long)nunits * split / nsplit), u_end = (int)((long long)nunits * (split + 1) / nsplit);
        if (tid == 0) *sCounter = u_begin;
        k2_trig_fill(a, sMu, sW, sTc, ldt, Kp, b0, tid, NC);
#pragma unroll 1
        for (int idx = tid; idx < 2 * K2_ANG * Q; idx += NC) {
            const int which = idx / (K2_ANG * Q), rem = idx - which * (K2_ANG * Q);
            const int ang = rem / Q, q = rem - ang * Q;
            long long b = b0 + ang; if (b >= a.nb) b = a.nb - 1;
            const double m = 0.5 * (double)(2 * q + (p.twoj & 1));
            double s, cs;
            sincos(m * (which ? a.gamma[b] : a.alpha[b]), &s, &cs);
            sPh[((2 * which + 0) * K2_ANG + ang) * Q + q] = cs;
            sPh[((2 * which + 1) * K2_ANG + ang) * Q + q] = s;
        }
        named_bar_sync(1, NC);

        // dynamic unit scheduler (one shared counter per chunk): unit v -> part v % upc of column v / upc: the eight
        // warps work on adjacent columns, whose output is adjacent in memory
#pragma unroll 1
        for (;;) {
            if (cx.bar_me) named_bar_sync(cx.bar_me, 64);
            int v = 0;
            if (lane == 0) v = atomicAdd(sCounter, 1);
            v = __shfl_sync(0xffffffffu, v, 0);
            if (v >= u_end) { if (cx.bar_other) named_bar_arrive(cx.bar_other, 64); break; }
            const int qn = v / upc, part = v - qn * upc;
            const int g0 = part * gbase + min(part, grem);
            const int gcnt = gbase + (part < grem ? 1 : 0);
            k2_unit_cplx<HALF>(a, cx, qn, g0, gcnt);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Real d, second-generation kernel (k2_real_kernel).  Same main loop; what changed around it:
//   * one staged tile feeds TWO mirror targets.  The butterfly sum S = P0 + P1 belongs to (i,i') and (ir,i'r), the
//     difference D = P0 - P1 to (ir,i') and (i,i'r): the compute warp stages S and D once each (half the STS of the
//     first-generation epilogue) and the store warp writes every staged value to an ascending and a descending
//     target (one LDS, two STG: half the LDS of the old drain);
//   * tiles are half as large (8 angles) and ride a 2-slot ring per compute warp: staging of tile t+1 overlaps
//     the drain of tile t, the compute warp is back in its main loop sooner;
//   * the NEXT chunk's trig tables are generated into a second buffer while the current chunk is computed (each
//     compute warp does its share after its first unit, while its partner warp keeps the DMMA pipe busy); the
//     per-chunk CTA barrier and the all-warps trig phase are gone (mbarriers trig_ready / trig_free instead);
//   * the last, partial wave of chunks is split over several CTAs each (K2Args::tsplit).
// ------------------------------------------------------------------------------------------------
constexpr int K2_TILE_ASC = 1, K2_TILE_DESC = 2, K2_TILE_END_ALL = 8;
constexpr int K2_TILE_MIRROR = 16;   // beta <-> pi - beta pairing: the tile also feeds the mirror matrices nb-1-b (ea2 / ed2)

struct K2Tile {                 // descriptor of one staged half tile: 8 angles (one MMA row tile) x <= 64 quadrant rows
    unsigned long long ea;      // ascending target : element index of (first angle, position 0); position s lives at ea + s
    unsigned long long ed;      // descending target: position s lives at ed - s
    int s_hi;                   // valid positions: s < s_hi (quadrant row < Q)
    int s_lo_d;                 // descending target only: s >= s_lo_d (quadrant row >= zskip)
    int nrows;                  // existing angles of the tile (1..8); angle ar of the tile is matrix 2*ar further on
    int dib_a, dib_d;           // (row - col) mod 4 of position 0 in the two targets: sigma(dib_a + s), sigma(dib_d - s)
    int kind;                   // K2_TILE_ASC | K2_TILE_DESC [| K2_TILE_MIRROR], or an END marker
    // pairing (d(pi-b)[m,n] = (-1)^(j+m) d(b)[m,-n], test/runtests.jl:330-334): the same staged values, with the row-alternating
    // extra sign, go to the mirror matrix of every angle -- ascending into the column that is the descending target's in the
    // front matrix and vice versa; angle ar of the tile is matrix 2*ar FURTHER BACK from ea2 / ed2
    unsigned long long ea2, ed2;
    int nrows2;                 // angles of the tile that have a mirror matrix (a prefix of nrows)
    int par;                    // bit 0: parity of the row of position 0 in the ascending target, bit 1: in the descending one
};
static_assert(sizeof(K2Tile) == 64, "K2Tile layout");

struct K2RSmem {                // offsets in doubles, computed identically on host and device
    int uq, mu, w, trig, stg, desc, bars, total;
    int ldt, qpad, trig_size;
};
constexpr int K2R_NBARS = 4 * K2_CWARPS + 6;   // full[8][2], empty[8][2], trig_ready[2], trig_free[2], UQ load, counters[2]
__host__ __device__ inline K2RSmem k2r_smem_layout(int Q, int Kp, int ldu)
{
    K2RSmem s;
    s.ldt = Kp | 4;
    s.qpad = (Q + 15) & ~15;
    s.trig_size = 2 * K2_ANG * s.ldt;                     // Tc[16][ldt] then Ts[16][ldt]
    int o = 0;
    s.uq = o;   o += s.qpad * ldu;
    s.mu = o;   o += Kp;
    s.w = o;    o += Kp;
    s.trig = o; o += 2 * s.trig_size;                     // two buffers: this chunk's tables and the next one's
    o = (o + 1) & ~1;
    s.stg = o;  o += K2_CWARPS * 2 * K2_HT_SIZE;
    s.desc = o; o += K2_CWARPS * 2 * (int)(sizeof(K2Tile) / 8);
    s.bars = o; o += K2R_NBARS;
    s.total = o;
    return s;
}

struct K2Item { long long b0; int u_begin, u_end; };
// work item -> (chunk of 16 angles, range of its units).  Items [0, nfull) are whole chunks; the remaining chunks are
// split tsplit ways each, so that the last wave (or a batch smaller than the GPU) still occupies every SM.
__device__ __forceinline__ K2Item k2_item(const K2Args &a, const long long it, const int nunits)
{
    K2Item o;
    if (it < a.nfull) { o.b0 = it * K2_ANG; o.u_begin = 0; o.u_end = nunits; return o; }
    const long long rr = it - a.nfull;
    const long long ch = a.nfull + rr / a.tsplit;
    const int split = (int)(rr - (rr / a.tsplit) * a.tsplit);
    o.b0 = ch * K2_ANG;
    o.u_begin = (int)((long long)nunits * split / a.tsplit);
    o.u_end = (int)((long long)nunits * (split + 1) / a.tsplit);
    return o;
}

struct K2RCtx {                 // what a compute warp needs for its hand-offs
    double *stg;                // 2 ring slots
    K2Tile *desc;               // 2 descriptors
    unsigned long long *full, *empty;   // 2 + 2 mbarriers
    unsigned tiles;             // tiles handed over so far
    size_t NN, e_item, e_ip, e_ipr;     // N*N; element index of the item's first matrix; of output columns i', i'r in it
    int ntiles, nrows0, nrows1; // existing angle tiles of the item and their angle counts (ragged last chunk)
    int paired, nrowsb0, nrowsb1;       // pairing: angles of the two tiles that have a mirror matrix
    size_t e_item_b;            // element index of the mirror matrix of the item's first angle
};

// Stage one butterfly result (S or D) of a unit, one angle tile at a time, and hand it to the store warp.
// rc.e_ip / rc.e_ipr: element index of (angle 0 of the chunk, row 0) of output columns i' and i'r (set once per unit).
template <int GMAX>
__device__ __forceinline__ void k2_emit(const K2Args &a, const K2Ctx &cx, K2RCtx &rc, const double (&val)[2][GMAX][2][2],
                                        const bool isD, const int qn, const int qbase, const int gcnt)
{
    const int lane = cx.lane;
    const bool both = qn >= cx.zskip;                               // n = 0 mirrors onto itself
    const int row_a = cx.i0 + qbase, row_d = cx.N - 1 - cx.i0 - qbase;
    const int ip = cx.i0 + qn;
#pragma unroll
    for (int ai = 0; ai < 2; ++ai) {
        if (ai >= rc.ntiles) break;                                 // ragged last chunk (warp-uniform)
        const unsigned t = rc.tiles;
        const int slot = t & 1;
        mbar_wait(rc.empty + slot, ((t >> 1) & 1) ^ 1);             // the store warp has drained tile t - 2
        if (lane == 0) {
            // S: (i,i') upwards, (ir,i'r) downwards;  D: (i,i'r) upwards, (ir,i') downwards
            K2Tile d;
            d.ea = (isD ? rc.e_ipr : rc.e_ip) + row_a + (ai ? rc.NN : 0);
            d.ed = (isD ? rc.e_ip : rc.e_ipr) + row_d + (ai ? rc.NN : 0);
            d.s_hi = min(16 * gcnt, cx.Q - qbase);
            d.s_lo_d = max(0, cx.zskip - qbase);
            d.nrows = ai ? rc.nrows1 : rc.nrows0;
            const int col_a = isD ? cx.N - 1 - ip : ip, col_d = isD ? ip : cx.N - 1 - ip;
            d.dib_a = (row_a - col_a) & 3; d.dib_d = (row_d - col_d) & 3;
            d.kind = a.nostore ? 0 : (isD ? (K2_TILE_DESC | (both ? K2_TILE_ASC : 0)) : (K2_TILE_ASC | (both ? K2_TILE_DESC : 0)));
            d.ea2 = d.ed2 = 0; d.nrows2 = 0; d.par = 0;
            if (rc.paired && d.kind) {
                // mirror matrix: S -> (i, i'r) upwards, (ir, i') downwards;  D -> (i, i') upwards, (ir, i'r) downwards
                const size_t eb = rc.e_item_b - (ai ? rc.NN : 0);
                const size_t c_ip = (size_t)ip * cx.N, c_ipr = (size_t)(cx.N - 1 - ip) * cx.N;
                d.ea2 = eb + (isD ? c_ip : c_ipr) + row_a;
                d.ed2 = eb + (isD ? c_ipr : c_ip) + row_d;
                d.nrows2 = ai ? rc.nrowsb1 : rc.nrowsb0;
                d.par = (row_a & 1) | ((row_d & 1) << 1);
                if (d.nrows2 > 0) d.kind |= K2_TILE_MIRROR;
            }
            rc.desc[slot] = d;
        }
        // the lane owns rows 16g+4c .. 16g+4c+3 of the unit: (even tile, odd tile) interleaved
        double *p = rc.stg + slot * K2_HT_SIZE + cx.r * K2_STG_LD + 4 * cx.c;
#pragma unroll
        for (int g = 0; g < GMAX; ++g) {
            if (g < gcnt) {
                reinterpret_cast<double2 *>(p + 16 * g)[0] = make_double2(val[ai][g][0][0], val[ai][g][1][0]);
                reinterpret_cast<double2 *>(p + 16 * g)[1] = make_double2(val[ai][g][0][1], val[ai][g][1][1]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(rc.full + slot);                 // release: tile + descriptor are visible
        rc.tiles = t + 1;
    }
}

__device__ __forceinline__ void k2_send_marker(const K2Ctx &cx, K2RCtx &rc, const int kind)
{
    const unsigned t = rc.tiles;
    const int slot = t & 1;
    mbar_wait(rc.empty + slot, ((t >> 1) & 1) ^ 1);
    if (cx.lane == 0) {
        K2Tile d;
        d.ea = d.ed = 0; d.s_hi = d.s_lo_d = d.nrows = d.dib_a = d.dib_d = 0; d.kind = kind;
        d.ea2 = d.ed2 = 0; d.nrows2 = 0; d.par = 0;
        rc.desc[slot] = d;
    }
    __syncwarp();
    if (cx.lane == 0) mbar_arrive(rc.full + slot);
    rc.tiles = t + 1;
}

// Store warp: one staged value -> its ascending and its descending target; 256 contiguous bytes per store instruction.
// All eight matrix pointers of a target are formed BEFORE its stores issue: a pointer that is advanced in place right
// after its store waits for the store to read it (write-after-read through the LSU queue, ~25 cycles per store in
// the ncu source view; the store warps were busy 83 % of the time and the compute warps waited for free slots).
// MIRROR: the targets of the mirror matrices (ea2 / ed2, matrices stepping backwards, row-alternating extra sign)
// (the descriptor stays in shared memory -- dp -- and is read field by field: the store warps run on 56 registers)
template <bool MIRROR>
__device__ __forceinline__ void k2_drain_targets(const double *buf, double *out, const K2Tile *dp, const size_t NN, const int lane)
{
    const long long step = MIRROR ? -2 * (long long)NN : 2 * (long long)NN;   // angle ar of the tile is matrix 2*ar further on (back)
    const int nrows = MIRROR ? dp->nrows2 : dp->nrows;
    const int s_hi = dp->s_hi, kind = dp->kind;
#pragma unroll 1
    for (int s = lane; s < s_hi; s += 32) {
        const double *sp = buf + s;
        const bool on_a = (kind & K2_TILE_ASC) != 0, on_d = (kind & K2_TILE_DESC) != 0 && s >= dp->s_lo_d;
        double *pa = out + (MIRROR ? dp->ea2 : dp->ea) + s, *pd = out + (MIRROR ? dp->ed2 : dp->ed) - s;
        unsigned sa = sigma_bit(dp->dib_a + s), sd = sigma_bit(dp->dib_d - s);
        if (MIRROR) { const int par = dp->par; sa ^= ((unsigned)(par + s) & 1u) << 31; sd ^= ((unsigned)((par >> 1) + s) & 1u) << 31; }   // (-1)^row
        if (nrows == K2_HT_ROWS) {                                    // all 8 loads in flight before the stores
            double v[K2_HT_ROWS];
            double *q[K2_HT_ROWS];
#pragma unroll
            for (int ar = 0; ar < K2_HT_ROWS; ++ar) v[ar] = sp[ar * K2_STG_LD];
            // signs are applied in place (one LOP3 on the high word per value and target, no register copies)
#pragma unroll
            for (int ar = 0; ar < K2_HT_ROWS; ++ar) v[ar] = flip(v[ar], sa);
            if (on_a) {
#pragma unroll
                for (int ar = 0; ar < K2_HT_ROWS; ++ar) q[ar] = pa + ar * step;
#pragma unroll
                for (int ar = 0; ar < K2_HT_ROWS; ++ar) *q[ar] = v[ar];
            }
            if (on_d) {
#pragma unroll
                for (int ar = 0; ar < K2_HT_ROWS; ++ar) { q[ar] = pd + ar * step; v[ar] = flip(v[ar], sa ^ sd); }
#pragma unroll
                for (int ar = 0; ar < K2_HT_ROWS; ++ar) *q[ar] = v[ar];
            }
        } else {
#pragma unroll 1
            for (int ar = 0; ar < nrows; ++ar) {
                const double v = sp[ar * K2_STG_LD];
                if (on_a) *pa = flip(v, sa);
                if (on_d) *pd = flip(v, sd);
                pa += step; pd += step;
            }
        }
    }
}
__device__ __forceinline__ void k2_drain_pair(const double *buf, double *out, const K2Tile *dp, const size_t NN, const int lane)
{
    k2_drain_targets<false>(buf, out, dp, NN, lane);
    if (dp->kind & K2_TILE_MIRROR) k2_drain_targets<true>(buf, out, dp, NN, lane);
}

// Store warp q: drains the 2-slot ring of compute warp q until the END marker.
//   stg : K2_CWARPS rings of 2 slots;  desc: K2_CWARPS x 2 descriptors;  bars: full[w][slot] then empty[w][slot]
__device__ __forceinline__ void k2_store_loop(const double *stg, const double *desc_base, unsigned long long *bars, const int q,
                                              double *out, const size_t NN, const int lane)
{
    const double *ring = stg + q * (2 * K2_HT_SIZE);
    const K2Tile *desc = reinterpret_cast<const K2Tile *>(desc_base) + 2 * q;
    unsigned long long *full = bars + 2 * q, *empty = bars + 2 * K2_CWARPS + 2 * q;
    for (unsigned t = 0;; ++t) {
        const int slot = t & 1;
        mbar_wait(full + slot, (t >> 1) & 1);                          // acquire: tile + descriptor are visible
        if (desc[slot].kind == K2_TILE_END_ALL) break;
        k2_drain_pair(ring + slot * K2_HT_SIZE, out, desc + slot, NN, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);                      // release: the slot may be overwritten
    }
}

template <bool HALF>
__device__ __forceinline__ void k2_unit_real(const K2Args &a, const K2Ctx &cx, K2RCtx &rc, const int qn, const int g0, const int gcnt)
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
    // butterfly: S = P0 + P1, D = P0 - P1 (half-integer j: D from the accumulators of the other trig function)
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
    rc.e_ip = rc.e_item + (size_t)(cx.i0 + qn) * cx.N;
    rc.e_ipr = rc.e_item + (size_t)(cx.N - 1 - cx.i0 - qn) * cx.N;
    k2_emit<GMAX>(a, cx, rc, vs, false, qn, 16 * g0, gcnt);
    k2_emit<GMAX>(a, cx, rc, vd, true, qn, 16 * g0, gcnt);
}

template <bool HALF>
__global__ void __launch_bounds__(K2_THREADS_REAL, 1) k2_real_kernel(K2Args a)
{
    constexpr int GMAX = HALF ? 2 : 4;
    constexpr int NT = K2_THREADS_REAL;
    constexpr int NC = K2_CWARPS * 32;
    extern __shared__ __align__(128) double sm[];
    // symmetric angle grids belong to k2_col_kernel (launched alongside): the device-side pair check decides, no host round trip
    if (a.skip_flag != nullptr && *reinterpret_cast<const volatile int *>(a.skip_flag) != 0) return;
    const PlanDev &p = a.p;
    const int Q = p.Q, Kp = p.Kp, ldu = p.ldu;
    const K2RSmem L = k2r_smem_layout(Q, Kp, ldu);
    const int ldt = L.ldt;
    double *sUQ = sm + L.uq, *sMu = sm + L.mu, *sW = sm + L.w;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + L.bars);
    unsigned long long *trig_ready = bars + 4 * K2_CWARPS, *trig_free = trig_ready + 2, *uq_bar = trig_free + 2;
    int *sCounter = reinterpret_cast<int *>(uq_bar + 1);                    // [2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int ngroups = L.qpad >> 4;
    const int upc = (ngroups + GMAX - 1) / GMAX;          // parts per column
    const int gbase = ngroups / upc, grem = ngroups - gbase * upc;   // the first grem parts have gbase+1 groups
    const int nunits = Q * upc;
    // this CTA's items: blockIdx.x, blockIdx.x + gridDim.x, ...  (the grid never exceeds the item count)
    const long long nmine = (a.nitems - blockIdx.x + gridDim.x - 1) / gridDim.x;

    // one-time: UQ -> shared through the TMA engine (one bulk copy, mbarrier-signalled); zero pad rows, mu, w by hand
    {
        const unsigned uq_bytes = (unsigned)(Q * ldu) * 8u;        // a multiple of 16: ldu is even
        if (tid == 0) {
            mbar_init(uq_bar, 1);
            for (int w = 0; w < 4 * K2_CWARPS; ++w) mbar_init(bars + w, 1);
            for (int b = 0; b < 2; ++b) { mbar_init(trig_ready + b, K2_CWARPS); mbar_init(trig_free + b, K2_CWARPS); }
            mbar_init_fence();
            mbar_expect_tx(uq_bar, uq_bytes);
            bulk_load(sUQ, p.uq, uq_bytes, uq_bar);
        }
        for (int i = Q * ldu + tid; i < L.qpad * ldu; i += NT) sUQ[i] = 0.0;
        for (int i = tid; i < Kp; i += NT) { sMu[i] = p.mu[i]; sW[i] = p.w[i]; }
        __syncthreads();                                           // the mbarriers are initialised for everybody
        mbar_wait(uq_bar, 0);
    }

    if (warp >= K2_CWARPS) {
        // ---- store warp q: drains the ring of compute warp q ----
        // (no trig work here: a call -- sincos' slow path -- below setmaxnreg.dec crashes ptxas 12.9)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(K2_REGS_STORE));
        k2_store_loop(sm + L.stg, sm + L.desc, bars, warp - K2_CWARPS, a.out, (size_t)p.N * p.N, lane);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(K2_REGS_COMPUTE));

    K2Ctx cx;
    cx.sUQ = sUQ; cx.sTc = nullptr; cx.sTs = nullptr; cx.sPh = nullptr;
    cx.stg = nullptr;
    cx.ldu = ldu; cx.ldt = ldt; cx.N = p.N; cx.Q = Q; cx.i0 = p.N - Q; cx.zskip = p.zskip;
    cx.Kp = Kp; cx.K0p = p.K0p; cx.twoj = p.twoj;
    cx.lane = lane; cx.r = lane >> 2; cx.c = lane & 3;
    K2RCtx rc;
    rc.stg = sm + L.stg + warp * (2 * K2_HT_SIZE);
    rc.desc = reinterpret_cast<K2Tile *>(sm + L.desc) + 2 * warp;
    rc.full = bars + 2 * warp; rc.empty = bars + 2 * K2_CWARPS + 2 * warp;
    rc.tiles = 0;
    rc.NN = (size_t)p.N * p.N;
    rc.paired = 0; rc.nrowsb0 = rc.nrowsb1 = 0; rc.e_item_b = 0;
    // hand-off between the two compute warps (w, w+4) of an SM sub-partition: see k2_cplx_kernel
    cx.krel = a.handoff - 1;
    cx.bar_me = a.handoff ? 2 + warp : 0;
    cx.bar_other = a.handoff ? 2 + (warp ^ 4) : 0;
    if (a.handoff && (warp & 4)) named_bar_arrive(cx.bar_other, 64);       // warps 0..3 take the first turn

    // The trig tables and the unit counter of item k+1 are made DURING item k, into the other buffer: every compute
    // warp contributes its share right after its first unit of item k -- its partner warp is then in its main loop,
    // so the DMMA pipe stays busy (the first-generation kernel stopped all warps at a CTA barrier for this).
    //   trig_ready[b]: 8 arrivals (one per compute warp) = the tables in buffer b are complete;
    //   trig_free[b] : 8 arrivals = every compute warp has finished the item that used buffer b.
    auto prefetch = [&](const long long kn) {
        const int bn = (int)(kn & 1);
        if (kn >= 2) mbar_wait(trig_free + bn, (unsigned)((kn - 2) >> 1) & 1u);      // item kn-2 has left this buffer
        const K2Item itn = k2_item(a, blockIdx.x + kn * gridDim.x, nunits);
        k2_trig_fill(a, sMu, sW, sm + L.trig + bn * L.trig_size, ldt, Kp, itn.b0, tid, NC);
        if (tid == 0) sCounter[bn] = itn.u_begin;
        __syncwarp();
        if (lane == 0) mbar_arrive(trig_ready + bn);
    };
    // (every sincos call site sits below setmaxnreg.inc: ptxas 12.9 crashes when the slow-path callee is reachable from
    // regions with different register budgets)
    prefetch(0);
#pragma unroll 1
    for (long long k = 0; k < nmine; ++k) {
        const int b = (int)(k & 1);
        const K2Item it = k2_item(a, blockIdx.x + k * gridDim.x, nunits);
        mbar_wait(trig_ready + b, (unsigned)(k >> 1) & 1u);           // this item's tables and counter are in place
        cx.sTc = sm + L.trig + b * L.trig_size; cx.sTs = cx.sTc + K2_ANG * ldt;
        cx.b0 = it.b0;
        rc.e_item = (size_t)it.b0 * rc.NN;
        rc.nrows0 = (int)min(8LL, (a.nb - it.b0 + 1) / 2);
        rc.nrows1 = (int)min(8LL, (a.nb - it.b0) / 2);
        rc.ntiles = rc.nrows1 > 0 ? 2 : 1;
        bool pre = k + 1 >= nmine;                                    // nothing left to prefetch
        // dynamic unit scheduler (one shared counter per item): unit v -> part v % upc of column v / upc.  The
        // eight warps then work on adjacent columns, whose output is adjacent in memory.
#pragma unroll 1
        for (;;) {
            // The unit index must be grabbed WHILE HOLDING the pair's token.  The token alternates strictly, every loop
            // iteration consumes it once and passes it on once; grabbing under the token makes "no units left" visible to
            // the two warps of a pair in alternation order, so that their iteration counts stay compatible.  (Grabbing
            // before the wait -- to hide the atomic's latency -- lets the second-turn warp win the last unit of a small
            // item while its partner has already left: its next wait for the token then never ends.  Measured: a hang.)
            if (cx.bar_me) named_bar_sync(cx.bar_me, 64);
            int v = 0;
            if (lane == 0) v = atomicAdd(sCounter + b, 1);
            v = __shfl_sync(0xffffffffu, v, 0);
            const bool have = v < it.u_end;
            if (have) {
                const int qn = v / upc, part = v - qn * upc;
                const int g0 = part * gbase + min(part, grem);
                const int gcnt = gbase + (part < grem ? 1 : 0);
                k2_unit_real<HALF>(a, cx, rc, qn, g0, gcnt);
            } else if (cx.bar_other) named_bar_arrive(cx.bar_other, 64);      // leaving: pass the token on
            if (!pre) { prefetch(k + 1); pre = true; }                        // after the first unit (or when there is none)
            if (!have) break;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(trig_free + b);                    // this warp is done with buffer b and its counter
    }
    k2_send_marker(cx, rc, K2_TILE_END_ALL);
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
    } else if (kind == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[i][0] = fma(acc[i][0], x, y); acc[i][1] = fma(acc[i][1], y, x); }
        }
    } else {
        // kind 2: full-range double sin (one evaluation per iteration), arguments spread over [0, 630) = mu * beta at j <= 200;
        // kind 3: one complex rotation step (4 multiply-adds), the recurrence that replaces most of the evaluations in k4_jmn_kernel
        double t = 0.1 + (blockIdx.x * blockDim.x + threadIdx.x) * 1e-3, c = 1.0, sn = 0.0;
        const double cd = cos(0.37), sd = sin(0.37);
        for (int it = 0; it < iters; ++it) {
            if (kind == 2) acc[0][0] += sin(t);
            else { const double c2 = c * cd - sn * sd; sn = fma(sn, cd, c * sd); c = c2; acc[0][0] = fma(c, t, acc[0][0]); }
            t += 1.618;
            if (t > 630.0) t -= 630.0;
        }
        acc[0][1] += c + sn;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) sink[0] = s;
}

}  // namespace wd
