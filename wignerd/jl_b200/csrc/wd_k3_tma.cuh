// wd_k3_tma.cuh -- second-generation complex-D kernel (k3_tma_kernel): the contraction of k2_col_kernel (one quadrant column x
// ALL quadrant rows x 8 angles per warp unit, DMMA.8x8x4) with the phase exp(-i(m alpha + n gamma)) of wignerD!
// (src/WignerD.jl:335-342) applied to the accumulators IN REGISTERS and the output leaving through the TMA engine as
// TENSOR stores (cp.async.bulk.tensor.2d, SASS UTMASTG): one instruction per box of [8 angles] x [half a column].
//
// Why (k2_cplx_kernel at cfg-3, j = 49.5, 10^5 triples: 3.96 ms = 62 % of the HBM write roof, 3.48 ms without its stores):
//   * its drain re-reads every staged value from shared memory, multiplies by four phase-table entries and issues two 16-byte
//     global stores per element: ~0.38 warp instructions per output element in a dependent chain, two warps per scheduler;
//   * complex elements are 16 bytes, so every matrix stride is a multiple of 16 bytes and the output is a legal TMA tensor
//     (dim0 = the 2 N^2 doubles of a matrix, dim1 = the batch): ONE instruction stores the half column of all 8 angles of an
//     item (k2_col_kernel's real output has odd N and needs 8 one-dimensional bulk copies per half column, ~65 cycles each);
//   * the lane that holds an accumulator also holds everything its phase needs: cos/sin(n gamma) of its angle (one pair per
//     unit) and cos/sin(m alpha) of its 4 consecutive rows (two 16-byte loads per table and 16-row group).  Per row and angle:
//     4 FP64 operations for (re, im) of the phase, 2 multiplies, and the value goes to its ascending target as (A, B) and to
//     its mirror target as +-(A, -B) (conjugate up to the sign sigma): 760 independent instructions per unit of 1600 outputs
//     with no global-memory instruction among them.
//
// Boxes of a unit (column n = qn >= 0, c0 = j+n, c1 = j-n; S = P0 + P1, D = P0 - P1; x = ca cg, y = sa sg, z = sa cg, w = ca sg):
//     c0 rows m >= 0  <- S (x - y, -(z + w))      c1 rows m <= 0  <- S (x - y, +(z + w))     (sign sigma(row - col) each)
//     c1 rows m >= 0  <- D (x + y,   w - z )      c0 rows m <= 0  <- D (x + y,   z - w )
// Staging: two box slots per warp ([8 angles][half column] complex, dense = the layout the tensor store reads), the store of
// box t-2 must have read its slot before box t is written (cp.async.bulk.wait_group.read 1).
// Trig and phase tables of item k+1 are made during item k (mbarriers, no CTA barrier), pair hand-off as in k2_col_kernel.
//
// Measured at cfg-3 (one B200): 3.19 ms = 5.01 TB/s = 77 % of the measured HBM write roof (k2_cplx_kernel: 3.96 ms); without
// its tensor stores ~3.1 ms, DMMA pipe 60 % (floor 1.94 ms): the SM side binds.  What was tried around it (a third staging
// box, 12 warps, software-pipelined epilogue pieces, merged accumulators, release points): DESIGN.md section 4.
// NW (warps per CTA) is a template parameter because the 12-warp form was measured; the library instantiates NW = 8 only.
#pragma once
#include <cuda.h>
#include "wd_k2_col.cuh"

namespace wd {

constexpr int K3T_ANG = 8;
constexpr int K3T_GMAX_INT = 4;             // integer j: Q <= 64 (more 16-row groups do not fit in 255 registers without spills)
constexpr int K3T_GMAX_HALF = 4;            // half-integer j (both trig functions accumulated): Q <= 64
constexpr int K3T_FLAG_NOSTORE = 1;         // experiment bit: no tensor stores
// Lanes c = 2, 3 of an MMA fragment row keep their rows pairwise swapped (row order 1,0,3,2 instead of 0,1,2,3): the 16-byte
// staging stores of a quarter warp (two angles x four lanes) then fall into eight different 16-byte bank groups when the box
// row is == 32 (mod 64) bytes long (j = 49.5: 800 bytes) instead of four (lanes c and c+2 are 128 bytes apart otherwise).
// The phase tables are stored in the same swapped order (position q ^ 1 for q in the upper half of a 16-row group), so the
// lane's 16-byte phase loads need no shuffling.
#ifndef K3T_SWAP
#define K3T_SWAP 1
#endif
__host__ __device__ inline int k3t_perm(int q) { return K3T_SWAP ? q ^ ((q >> 3) & 1) : q; }

struct K3TArgs {
    PlanDev p;
    const double *alpha, *beta, *gamma;
    long long nb;
    int handoff;              // see K2Args::handoff
    int flags;
};

struct K3TSmem {                // offsets in doubles, identical on host and device
    int uq, mu, w, trig, ph, stg, bars, total;
    int ldt, qp, trig_size, ph_size, slot;
};
__host__ __device__ inline K3TSmem k3t_smem_layout(int Q, int Kp, int ldu, int nw)
{
    K3TSmem s;
    s.ldt = Kp | 4;
    // phase-table row: every row a lane can touch (16 GC) is in bounds; == 2 (mod 4) doubles, so that the 16-byte loads of the
    // two angles of a quarter warp interleave instead of colliding
    s.qp = 16 * ((Q + 15) >> 4) + 2;
    s.trig_size = 2 * K3T_ANG * s.ldt;
    s.ph_size = 4 * K3T_ANG * s.qp;                  // cos(m alpha), sin(m alpha), cos(n gamma), sin(n gamma): [8][qp] each
    s.slot = K3T_ANG * 2 * Q;                        // one box: 8 angles x Q complex numbers (a multiple of 128 bytes)
    int o = 0;
    s.uq = o;    o += Q * ldu;
    s.mu = o;    o += Kp;
    s.w = o;     o += Kp;
    s.trig = o;  o += 2 * s.trig_size;
    o = (o + 1) & ~1;
    s.ph = o;    o += 2 * s.ph_size;
    o = (o + 15) & ~15;                              // boxes start on 128-byte boundaries
    s.stg = o;   o += nw * 2 * s.slot;
    s.bars = o;  o += 8;                             // trig_ready[2], trig_free[2], UQ load
    s.total = o;
    return s;
}
__host__ __device__ inline bool k3t_shape_ok(int twoj)
{
    const int Q = (twoj + 2) / 2;
    return twoj >= 5 && ((Q + 15) >> 4) <= ((twoj & 1) ? K3T_GMAX_HALF : K3T_GMAX_INT);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, const void *ssrc, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((unsigned long long)tm),
                 "r"(smem_addr(ssrc)), "r"(c0), "r"(c1)
                 : "memory");
}

struct K3TWarp {                // per-warp state of the store path
    double *stg;                // two box slots
    const double *ph;           // phase tables of the current item
    const CUtensorMap *tm_asc, *tm_desc;
    int N, Q, i0, zskip, qp, slot;
    int lane, r, c;
    int nvl;                    // valid rows (0..4) of this lane in the LAST 16-row group of a column
    int f0;                     // first angle of the item
    unsigned steps;             // boxes so far (slot = steps & 1)
    int nostore;
};

// One box: the lane's phase-applied values (A, B) of rows 16 g + 4 c + jj go to the staging slot in the order the output
// column has them -- ascending part: position q; descending part (rows m <= 0): position Q-1-q, value eps (A, -B) -- and one
// lane sends the box off.
template <int GC, bool DESC>
__device__ __forceinline__ void k3t_box(K3TWarp &ws, const double (&A)[GC][4], const double (&B)[GC][4], const unsigned (&me)[4], const int col)
{
    const int Q = ws.Q;
    double *slot = ws.stg + (ws.steps & 1u) * ws.slot;
    bulk_wait_read<1>();                                              // the store that read this slot two boxes ago is done
    __syncwarp();
    {
        const int rs = 2 * (DESC ? ws.i0 : Q);                        // dense box rows
        const int swb = (K3T_SWAP && (ws.c & 2)) ? 1 : 0;             // register jj holds row 16 g + 4 c + (jj ^ swb)
        double *base = slot + ws.r * rs + (DESC ? 2 * (Q - 1 - 4 * ws.c) : 8 * ws.c);
        double *base_e = DESC ? base - 2 * swb : base + 2 * swb;      // even jj: row jj + swb; odd jj: row jj - swb
        double *base_o = DESC ? base + 2 * swb : base - 2 * swb;
        const bool skip0 = DESC && ws.zskip && ws.c == 0;             // q = 0 is m = 0: it mirrors onto itself
#pragma unroll
        for (int g = 0; g < GC; ++g) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const bool ok = (g < GC - 1 || (jj ^ swb) < ws.nvl) && !(g == 0 && jj == 0 && skip0);      // only the last group is partial
                double *bp = (jj & 1) ? base_o : base_e;
                if (ok) {
                    if (DESC) *reinterpret_cast<double2 *>(bp - 2 * (16 * g + jj)) = make_double2(flip(A[g][jj], me[jj]), flip(B[g][jj], me[jj] ^ 0x80000000u));
                    else      *reinterpret_cast<double2 *>(bp + 2 * (16 * g + jj)) = make_double2(A[g][jj], B[g][jj]);
                }
            }
        }
    }
    fence_async_smem();
    __syncwarp();
    if (ws.lane == 0 && !ws.nostore) tma_store_2d(DESC ? ws.tm_desc : ws.tm_asc, slot, 2 * (col * ws.N + (DESC ? 0 : ws.i0)), ws.f0);
    bulk_commit();
    ws.steps++;
}

// The two boxes that one butterfly result feeds (ISD = false: S -> c0 ascending, c1 descending; true: D -> c1 ascending, c0 descending)
template <int GC, bool ISD>
__device__ __forceinline__ void k3t_emit(K3TWarp &ws, const double (&v)[GC][4], const int qn)
{
    const int qp = ws.qp, r = ws.r, c = ws.c;
    const double cg = ws.ph[(2 * K3T_ANG + r) * qp + k3t_perm(qn)];
    const double sg0 = ws.ph[(3 * K3T_ANG + r) * qp + k3t_perm(qn)];
    const double sgx = ISD ? sg0 : -sg0;
    // row - column of the ascending target for the lane's row jj = 0 of a group (16 g + 4 c vanish mod 4): S: q - qn; D: q + qn + 1 - zskip.
    // The descending target has the opposite difference: sigma(-d) = sigma(d) (-1)^d.
    const int d0 = ISD ? qn + 1 - ws.zskip : -qn;
    unsigned ms[4], me[4];
    const int swb = (K3T_SWAP && (c & 2)) ? 1 : 0;                    // register jj holds row 16 g + 4 c + (jj ^ swb)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) { ms[jj] = sigma_bit(d0 + (jj ^ swb)); me[jj] = ((unsigned)(d0 + (jj ^ swb)) & 1u) << 31; }
    double A[GC][4], B[GC][4];
    const double *pca = ws.ph + r * qp + 4 * c, *psa = pca + K3T_ANG * qp;
#pragma unroll
    for (int g = 0; g < GC; ++g) {
        const double2 c01 = *reinterpret_cast<const double2 *>(pca + 16 * g), c23 = *reinterpret_cast<const double2 *>(pca + 16 * g + 2);
        const double2 s01 = *reinterpret_cast<const double2 *>(psa + 16 * g), s23 = *reinterpret_cast<const double2 *>(psa + 16 * g + 2);
        const double ca[4] = {c01.x, c01.y, c23.x, c23.y}, sa[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            // S: (x - y, -(z + w));  D: (x + y, w - z)  -- the same two fused operations with sgx = -+ sin(n gamma)
            const double y = sa[jj] * sgx, z = sa[jj] * cg;
            const double re = fma(ca[jj], cg, y), im = fma(ca[jj], sgx, -z);
            const double vv = flip(v[g][jj], ms[jj]);
            A[g][jj] = vv * re;
            B[g][jj] = vv * im;
        }
    }
    const int c0 = ws.i0 + qn, c1 = ws.Q - 1 - qn;
    const bool mirror = qn >= ws.zskip;                               // n = 0 mirrors onto itself: no column c1
    if (!ISD || mirror) k3t_box<GC, false>(ws, A, B, me, ISD ? c1 : c0);
    if (ISD || mirror) k3t_box<GC, true>(ws, A, B, me, ISD ? c0 : c1);
}

// One unit: quadrant column qn x all quadrant rows (GC groups of 16) x 8 angles.  The main loop IS k2_col_kernel's
// (k2c_mainloop): the same DMMAs in the same order as the real-d kernels, so that wignerD(j, 0, beta, 0) equals wignerd(j, beta)
// bit for bit, which the reference's tests demand (test/runtests.jl:343).  (Tried: for half-integer j, both eigenvector classes
// accumulated into ONE accumulator set per trig function -- S = main function, D = other function with the A operand of class 1
// negated -- no butterfly, 32 instead of 64 accumulator doubles: 3.22 instead of 3.25 ms at cfg-3, but the sums round
// differently from the real-d path and that equality is lost.)
template <bool HALF, int GC>
__device__ __forceinline__ void k3t_unit(const K2Ctx &cx, K3TWarp &ws, const int qn)
{
    constexpr int NALT = HALF ? 2 : 1;
    double acc[2][GC][2][NALT][2];
#pragma unroll
    for (int i = 0; i < 2 * GC * 2 * NALT * 2; ++i) (&acc[0][0][0][0][0])[i] = 0.0;
    k2c_mainloop<HALF, GC>(acc, cx, qn);
    // butterfly: S = P0 + P1, D = P0 - P1 (half-integer j: D from the other trig function).  [group][register jj]: row
    // 16 g + 4 c + (jj ^ swb), jj = 2 h + eo (lanes c = 2, 3 keep their row pairs swapped: K3T_SWAP)
    double vS[GC][4], vD[GC][4];
    const bool sw = K3T_SWAP && (ws.c & 2);
#pragma unroll
    for (int g = 0; g < GC; ++g)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double s0 = acc[0][g][0][0][h] + acc[1][g][0][0][h], s1 = acc[0][g][1][0][h] + acc[1][g][1][0][h];
            const double d0 = acc[0][g][0][NALT - 1][h] - acc[1][g][0][NALT - 1][h], d1 = acc[0][g][1][NALT - 1][h] - acc[1][g][1][NALT - 1][h];
            vS[g][2 * h] = sw ? s1 : s0; vS[g][2 * h + 1] = sw ? s0 : s1;
            vD[g][2 * h] = sw ? d1 : d0; vD[g][2 * h + 1] = sw ? d0 : d1;
        }
    k3t_emit<GC, false>(ws, vS, qn);
    k3t_emit<GC, true>(ws, vD, qn);
}

template <bool HALF, int GC, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k3_tma_kernel(const K3TArgs a, const __grid_constant__ CUtensorMap tm_asc,
                                                                 const __grid_constant__ CUtensorMap tm_desc)
{
    extern __shared__ __align__(128) double sm[];
    const PlanDev &p = a.p;
    const int Q = p.Q, Kp = p.Kp, ldu = p.ldu, N = p.N;
    const K3TSmem L = k3t_smem_layout(Q, Kp, ldu, NW);
    const int ldt = L.ldt, qp = L.qp;
    double *sUQ = sm + L.uq, *sMu = sm + L.mu, *sW = sm + L.w;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + L.bars);
    unsigned long long *trig_ready = bars, *trig_free = bars + 2, *uq_bar = bars + 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // items: chunks of 8 angles; the last, partial wave is split over the idle CTAs (column ranges)
    const long long nch = (a.nb + K3T_ANG - 1) / K3T_ANG, G = gridDim.x;
    const long long nfull = (nch / G) * G, rem = nch - nfull;
    const int tsplit = rem ? (int)max(1LL, min(16LL, G / rem)) : 1;
    const long long nitems = nfull + rem * tsplit;
    if ((long long)blockIdx.x >= nitems) return;
    const long long nmine = (nitems - blockIdx.x + G - 1) / G;
    auto item_of = [&](const long long it) {
        K2CItem o;
        long long ch;
        if (it < nfull) { ch = it; o.q_begin = 0; o.q_end = Q; }
        else {
            const long long rr = it - nfull;
            ch = nfull + rr / tsplit;
            const int split = (int)(rr - (rr / tsplit) * tsplit);
            o.q_begin = (int)((long long)Q * split / tsplit);
            o.q_end = (int)((long long)Q * (split + 1) / tsplit);
        }
        o.f0 = ch * K3T_ANG;
        o.na = (int)min((long long)K3T_ANG, a.nb - o.f0);
        return o;
    };

    {
        const unsigned uq_bytes = (unsigned)(Q * ldu) * 8u;        // a multiple of 16: ldu is even
        if (tid == 0) {
            mbar_init(uq_bar, 1);
            for (int b = 0; b < 2; ++b) { mbar_init(trig_ready + b, NW); mbar_init(trig_free + b, NW); }
            mbar_init_fence();
            mbar_expect_tx(uq_bar, uq_bytes);
            bulk_load(sUQ, p.uq, uq_bytes, uq_bar);
        }
        for (int i = tid; i < Kp; i += (NW * 32)) { sMu[i] = p.mu[i]; sW[i] = p.w[i]; }
        __syncthreads();
        mbar_wait(uq_bar, 0);
    }

    K2Ctx cx;
    cx.sUQ = sUQ; cx.sTc = nullptr; cx.sTs = nullptr; cx.sPh = nullptr; cx.stg = nullptr;
    cx.ldu = ldu; cx.ldt = ldt; cx.N = N; cx.Q = Q; cx.i0 = N - Q; cx.zskip = p.zskip;
    cx.Kp = Kp; cx.K0p = p.K0p; cx.twoj = p.twoj;
    cx.lane = lane; cx.r = lane >> 2; cx.c = lane & 3;
    cx.b0 = 0;
    cx.krel = a.handoff - 1;
    // the warps w, w + 4, (w + 8) share an SM sub-partition and hence one DMMA pipe: a token goes round (named barriers), a warp
    // starts a main loop only when it holds it and passes it on at its release point (see k2_cplx_kernel)
    cx.bar_me = a.handoff ? 2 + warp : 0;
    cx.bar_other = a.handoff ? 2 + (warp + 4) % NW : 0;
    if (a.handoff && warp >= NW - 4) named_bar_arrive(cx.bar_other, 64);   // warps 0..3 take the first turn

    K3TWarp ws;
    ws.stg = sm + L.stg + warp * (2 * L.slot);
    ws.ph = nullptr;
    ws.tm_asc = &tm_asc; ws.tm_desc = &tm_desc;
    ws.N = N; ws.Q = Q; ws.i0 = N - Q; ws.zskip = p.zskip; ws.qp = qp; ws.slot = L.slot;
    ws.lane = lane; ws.r = lane >> 2; ws.c = lane & 3;
    ws.nvl = max(0, min(4, Q - 16 * (GC - 1) - 4 * (lane & 3)));
    ws.steps = 0; ws.f0 = 0;
    ws.nostore = (a.flags & K3T_FLAG_NOSTORE) != 0;

    // tables of item k+1 are made during item k: Tc/Ts[angle][kp] = w cos|sin(mu beta) (fl(mu*beta) then sincos like
    // src/WignerD.jl:199) and the phase tables cos/sin(m alpha), cos/sin(n gamma) for m, n = the quadrant's non-negative values
    const double mhalf = 0.5 * (double)(p.twoj & 1);
    auto prefetch = [&](const long long kn) {
        const int bn = (int)(kn & 1);
        if (kn >= 2) mbar_wait(trig_free + bn, (unsigned)((kn - 2) >> 1) & 1u);
        const K2CItem itn = item_of(blockIdx.x + kn * G);
        double *tc = sm + L.trig + bn * L.trig_size, *ts = tc + K3T_ANG * ldt;
#pragma unroll 1
        for (int idx = tid; idx < K3T_ANG * Kp; idx += (NW * 32)) {
            const int ang = idx / Kp, kp = idx - ang * Kp;
            long long b = itn.f0 + ang; if (b >= a.nb) b = a.nb - 1;
            double sn, cs;
            sincos(sMu[kp] * a.beta[b], &sn, &cs);
            const double w = sW[kp];
            tc[ang * ldt + kp] = w * cs;
            ts[ang * ldt + kp] = w * sn;
        }
        // phase tables: runs of 8 consecutive m (or n) per thread -- an exact sincos(fl(m x)) anchor at the run's first entry, then
        // rotations by x (<= 7 steps: the trig values stay within ~1e-15 of the anchored form; one sincos per entry made the
        // tables 10 % of a warp's time at cfg-3)
        double *ph = sm + L.ph + bn * L.ph_size;
        const int nruns = (Q + 7) >> 3;
#pragma unroll 1
        for (int idx = tid; idx < 2 * K3T_ANG * nruns; idx += (NW * 32)) {
            const int which = idx / (K3T_ANG * nruns), rem2 = idx - which * (K3T_ANG * nruns);
            const int ang = rem2 / nruns, q0 = 8 * (rem2 - ang * nruns);
            long long b = itn.f0 + ang; if (b >= a.nb) b = a.nb - 1;
            const double x = which ? a.gamma[b] : a.alpha[b];
            double sx, cx1, sn, cs;
            sincos(x, &sx, &cx1);
            sincos(((double)q0 + mhalf) * x, &sn, &cs);
            double *pc = ph + ((2 * which + 0) * K3T_ANG + ang) * qp, *ps = ph + ((2 * which + 1) * K3T_ANG + ang) * qp;
            const int qe = min(q0 + 8, Q);
#pragma unroll 1
            for (int q = q0; q < qe; ++q) {
                pc[k3t_perm(q)] = cs;
                ps[k3t_perm(q)] = sn;
                const double c2 = fma(cs, cx1, -(sn * sx)), s2 = fma(sn, cx1, cs * sx);
                cs = c2; sn = s2;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(trig_ready + bn);
    };
    prefetch(0);
#pragma unroll 1
    for (long long k = 0; k < nmine; ++k) {
        const int b = (int)(k & 1);
        const K2CItem it = item_of(blockIdx.x + k * G);
        mbar_wait(trig_ready + b, (unsigned)(k >> 1) & 1u);
        cx.sTc = sm + L.trig + b * L.trig_size; cx.sTs = cx.sTc + K3T_ANG * ldt;
        ws.ph = sm + L.ph + b * L.ph_size;
        ws.f0 = (int)it.f0;
        // the warps work on adjacent columns (rotated from item to item so that the uneven shares even out)
        const int ncols = it.q_end - it.q_begin, wslot = (warp + (int)(k % NW)) % NW;
        const int iters = (ncols + NW - 1) / NW;        // the warps of a ring make the same number of turns
        bool pre = k + 1 >= nmine;
#pragma unroll 1
        for (int u = 0; u < iters; ++u) {
            const int qn = it.q_begin + u * NW + wslot;
            if (cx.bar_me) named_bar_sync(cx.bar_me, 64);
            if (qn < it.q_end) k3t_unit<HALF, GC>(cx, ws, qn);
            else if (cx.bar_other) named_bar_arrive(cx.bar_other, 64);        // idle turn: pass the token on
            if (!pre) { prefetch(k + 1); pre = true; }
        }
        if (!pre) prefetch(k + 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(trig_free + b);
    }
    bulk_wait_all();
}

}  // namespace wd
