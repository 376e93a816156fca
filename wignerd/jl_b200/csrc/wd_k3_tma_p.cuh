// wd_k3_tma_p.cuh -- k3_tmap_kernel: the complex-D tensor-store kernel (wd_k3_tma.cuh) with THREE warps per SM sub-partition.
//
// k3_tma_kernel is SM-bound (cfg-3: 3.14 ms without its stores, DMMA floor 1.94 ms): a single warp's main loop keeps the DMMA
// pipe only ~65 % busy (in-order issue: the non-DMMA instructions of a k-step are not hidden behind the warp's own DMMAs), and
// with two warps per sub-partition one of them is in its epilogue most of the time.  Here TWO main loops run beside each
// other while the third warp of the sub-partition is in its epilogue:
//   * 12 warps need <= 168 registers: a unit is one quadrant column x TWO 16-row groups (32 rows) x 8 angles -- 32 accumulator
//     doubles even for half-integer j -- so a column is cut into <= 2 parts and a box is [8 angles] x [a quarter column]
//     (four tensor maps: ascending / descending part 0 / 1);
//   * the ring w -> w+4 -> w+8 -> w of a sub-partition carries two tokens; a warp starts a main loop when it holds one and
//     passes it on at its release point.  Tokens are counting semaphores in shared memory (one producer, one consumer each:
//     a monotonic counter written with a volatile store, polled with nanosleep) -- a named barrier cannot hold two arrivals.
// Everything else (phase in registers, row-pair swap, staging ring, tables one item ahead) is k3_tma_kernel's.
#pragma once
#include "wd_k3_tma.cuh"

namespace wd {

constexpr int K3P_WARPS = 12;
constexpr int K3P_THREADS = K3P_WARPS * 32;
constexpr int K3P_GP = 2;                   // 16-row groups per unit
constexpr int K3P_ROWS = 16 * K3P_GP;       // rows per part
constexpr int K3P_SLOT = K3T_ANG * 2 * K3P_ROWS;   // doubles per staging box (8 angles x 32 complex = 4 KB)

struct K3PArgs {
    PlanDev p;
    const double *alpha, *beta, *gamma;
    long long nb;
    int handoff;              // release point of the token, K2Args::handoff encoding; 0 = no tokens
    int flags;
    int nslots;               // staging boxes per warp (2 or 3)
};

struct K3PSmem {                // offsets in doubles, identical on host and device
    int uq, mu, w, trig, ph, stg, bars, tok, total;
    int ldt, qp, trig_size, ph_size;
};
__host__ __device__ inline K3PSmem k3p_smem_layout(int Q, int Kp, int ldu, int nslots)
{
    K3PSmem s;
    s.ldt = Kp | 4;
    // phase-table row: see k3t_smem_layout; a unit always touches two 16-row groups, so whole parts are in bounds
    s.qp = K3P_ROWS * ((((Q + 15) >> 4) + K3P_GP - 1) / K3P_GP) + 2;
    s.trig_size = 2 * K3T_ANG * s.ldt;
    s.ph_size = 4 * K3T_ANG * s.qp;
    int o = 0;
    s.uq = o;    o += Q * ldu;
    s.mu = o;    o += Kp;
    s.w = o;     o += Kp;
    s.trig = o;  o += 2 * s.trig_size;
    o = (o + 1) & ~1;
    s.ph = o;    o += 2 * s.ph_size;
    o = (o + 15) & ~15;                              // boxes start on 128-byte boundaries
    s.stg = o;   o += K3P_WARPS * nslots * K3P_SLOT;
    s.bars = o;  o += 8;                             // trig_ready[2], trig_free[2], UQ load
    s.tok = o;   o += (K3P_WARPS + 1) / 2;           // int tok[12]
    s.total = o;
    return s;
}

// box geometry of part `part` (rows 32 part .. of the quadrant) -- shared by the host (tensor maps) and the kernel
struct K3PBox { int wa, wd, p0; };                   // ascending width, descending width, first descending position (complex elements)
__host__ __device__ inline K3PBox k3p_box(int Q, int zskip, int part)
{
    K3PBox b;
    const int q0 = K3P_ROWS * part;
    b.wa = Q - q0 < K3P_ROWS ? Q - q0 : K3P_ROWS;
    b.p0 = Q - (q0 + K3P_ROWS) > 0 ? Q - (q0 + K3P_ROWS) : 0;
    b.wd = (Q - 1 - q0 - (part == 0 ? zskip : 0)) - b.p0 + 1;
    return b;
}

struct K3PMaps { CUtensorMap asc[2], desc[2]; };

struct K3PWarp {
    double *stg;
    const double *ph;
    const K3PMaps *tm;
    volatile int *tok_me, *tok_next;
    int tok_init_next, consumed, released;
    int N, Q, i0, zskip, qp;
    int lane, r, c;
    int f0, nslots, cur, nostore;
};

__device__ __forceinline__ void k3p_acquire(K3PWarp &ws)
{
    if (!ws.tok_me) return;
    if (ws.lane == 0) { while (*ws.tok_me <= ws.consumed) __nanosleep(40); }
    ws.consumed++;
    __syncwarp();
}
__device__ __forceinline__ void k3p_release(K3PWarp &ws)
{
    if (!ws.tok_next) return;
    __syncwarp();
    ws.released++;
    if (ws.lane == 0) *ws.tok_next = ws.tok_init_next + ws.released;
}

// main loop of one unit: column qn, rows 16 (g0 + g) + .. for g = 0, 1, 8 angles; the DMMA sequence of every accumulator is the
// real-d kernels' (k2c_mainloop).  Rows past the last one are clamped (computed on duplicate data, never stored).
template <bool HALF>
__device__ __forceinline__ void k3p_mainloop(double (&acc)[2][K3P_GP][2][HALF ? 2 : 1][2], const K2Ctx &cx, K3PWarp &ws, const int qn, const int g0)
{
    const int ldu = cx.ldu, ldt = cx.ldt, r = cx.r, c = cx.c, Q = cx.Q;
    const int pe = qn & 1;
    const double *tA = (pe ? cx.sTs : cx.sTc) + r * ldt + c;      // trig for the even-row tiles
    const double *tB = (pe ? cx.sTc : cx.sTs) + r * ldt + c;      // trig for the odd-row tiles
    const double *un = cx.sUQ + qn * ldu + c;
    const double *ue[K3P_GP], *uo[K3P_GP];
#pragma unroll
    for (int g = 0; g < K3P_GP; ++g) {
        ue[g] = cx.sUQ + min(16 * (g0 + g) + 2 * r, Q - 1) * ldu + c;
        uo[g] = cx.sUQ + min(16 * (g0 + g) + 2 * r + 1, Q - 1) * ldu + c;
    }
#pragma unroll
    for (int cls = 0; cls < 2; ++cls) {
        const int kbeg = cls ? cx.K0p : 0;
        const int kend = cls ? cx.Kp : cx.K0p;
        struct Frag { double un, ta, tb, be[K3P_GP], bo[K3P_GP]; };
        auto load = [&](Frag &f, const int k) {
            f.un = un[k]; f.ta = tA[k]; f.tb = tB[k];
#pragma unroll
            for (int g = 0; g < K3P_GP; ++g) { f.be[g] = ue[g][k]; f.bo[g] = uo[g][k]; }
        };
        auto step = [&](Frag &f) {
            f.ta *= f.un; f.tb *= f.un;                               // A = T(beta, kp) * UQ[qn][kp]
#pragma unroll
            for (int g = 0; g < K3P_GP; ++g) {
                dmma(acc[cls][g][0][0], f.ta, f.be[g]);
                dmma(acc[cls][g][1][0], f.tb, f.bo[g]);
                if (HALF) {
                    dmma(acc[cls][g][0][1], f.tb, f.be[g]);
                    dmma(acc[cls][g][1][1], f.ta, f.bo[g]);
                }
            }
        };
        Frag f0, f1;
        load(f0, kbeg);
        int k = kbeg;
        const int krel = cls == 0 ? (cx.krel < 100 ? kbeg + 8 * cx.krel : -1) : (cx.krel >= 100 ? kbeg + 8 * (cx.krel - 100) : -1);
        bool released = false;
#pragma unroll 1
        for (; k + 4 < kend; k += 8) {
            if (k == krel) { k3p_release(ws); released = true; }
            load(f1, k + 4);
            step(f0);
            load(f0, k + 8);          // past the last step: reads (and discards) in-bounds shared memory
            step(f1);
        }
        if (k < kend) step(f0);
        if (!released && ((cls == 0 && cx.krel < 100) || (cls == 1 && cx.krel >= 100))) k3p_release(ws);
    }
}

// one box of a unit: see k3t_box.  pos0: position of row q = 16 g0 inside the box (ascending) / of row q = 16 g0 counted
// downwards (descending); wbox: box width in complex elements (dense rows)
template <bool DESC>
__device__ __forceinline__ void k3p_boxout(K3PWarp &ws, const double (&A)[K3P_GP][4], const double (&B)[K3P_GP][4], const unsigned (&me)[4],
                                           const int g0, const int wbox, const int pos0, const CUtensorMap *tm, const int coord0)
{
    double *slot = ws.stg + ws.cur * K3P_SLOT;
    if (ws.nslots == 3) bulk_wait_read<2>(); else bulk_wait_read<1>();    // the store that read this slot nslots boxes ago is done
    __syncwarp();
    {
        const int swb = (K3T_SWAP && (ws.c & 2)) ? 1 : 0;             // register jj holds row 16 (g0 + g) + 4 c + (jj ^ swb)
        double *row = slot + ws.r * (2 * wbox);
#pragma unroll
        for (int g = 0; g < K3P_GP; ++g) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int ql = 16 * g + 4 * ws.c + (jj ^ swb);        // row within the part
                const int q = 16 * g0 + ql;
                const bool ok = q < ws.Q && !(DESC && ws.zskip && q == 0);      // q = 0 is m = 0: it mirrors onto itself
                if (ok) {
                    if (DESC) *reinterpret_cast<double2 *>(row + 2 * (pos0 - ql)) = make_double2(flip(A[g][jj], me[jj]), flip(B[g][jj], me[jj] ^ 0x80000000u));
                    else      *reinterpret_cast<double2 *>(row + 2 * (pos0 + ql)) = make_double2(A[g][jj], B[g][jj]);
                }
            }
        }
    }
    fence_async_smem();
    __syncwarp();
    if (ws.lane == 0 && !ws.nostore) tma_store_2d(tm, slot, coord0, ws.f0);
    bulk_commit();
    ws.cur = ws.cur + 1 == ws.nslots ? 0 : ws.cur + 1;
}

template <bool ISD>
__device__ __forceinline__ void k3p_emit(K3PWarp &ws, const double (&v)[K3P_GP][4], const int qn, const int part)
{
    const int qp = ws.qp, r = ws.r, c = ws.c, Q = ws.Q, g0 = K3P_GP * part;
    const double cg = ws.ph[(2 * K3T_ANG + r) * qp + k3t_perm(qn)];
    const double sg0 = ws.ph[(3 * K3T_ANG + r) * qp + k3t_perm(qn)];
    const double sgx = ISD ? sg0 : -sg0;
    const int d0 = ISD ? qn + 1 - ws.zskip : -qn;                     // see k3t_emit (16 g0 + 16 g + 4 c vanish mod 4)
    const int swb = (K3T_SWAP && (c & 2)) ? 1 : 0;
    unsigned ms[4], me[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) { ms[jj] = sigma_bit(d0 + (jj ^ swb)); me[jj] = ((unsigned)(d0 + (jj ^ swb)) & 1u) << 31; }
    double A[K3P_GP][4], B[K3P_GP][4];
    const double *pca = ws.ph + r * qp + 16 * g0 + 4 * c, *psa = pca + K3T_ANG * qp;
#pragma unroll
    for (int g = 0; g < K3P_GP; ++g) {
        const double2 c01 = *reinterpret_cast<const double2 *>(pca + 16 * g), c23 = *reinterpret_cast<const double2 *>(pca + 16 * g + 2);
        const double2 s01 = *reinterpret_cast<const double2 *>(psa + 16 * g), s23 = *reinterpret_cast<const double2 *>(psa + 16 * g + 2);
        const double ca[4] = {c01.x, c01.y, c23.x, c23.y}, sa[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const double y = sa[jj] * sgx, z = sa[jj] * cg;
            const double re = fma(ca[jj], cg, y), im = fma(ca[jj], sgx, -z);
            const double vv = flip(v[g][jj], ms[jj]);
            A[g][jj] = vv * re;
            B[g][jj] = vv * im;
        }
    }
    const int c0 = ws.i0 + qn, c1 = Q - 1 - qn;
    const bool mirror = qn >= ws.zskip;                               // n = 0 mirrors onto itself: no column c1
    const K3PBox bx = k3p_box(Q, ws.zskip, part);
    const int q0 = 16 * g0;
    if (!ISD || mirror) k3p_boxout<false>(ws, A, B, me, g0, bx.wa, 0, &ws.tm->asc[part], 2 * ((ISD ? c1 : c0) * ws.N + ws.i0 + q0));
    if ((ISD || mirror) && bx.wd > 0) k3p_boxout<true>(ws, A, B, me, g0, bx.wd, (Q - 1 - q0) - bx.p0, &ws.tm->desc[part], 2 * ((ISD ? c0 : c1) * ws.N + bx.p0));
}

template <bool HALF>
__device__ __forceinline__ void k3p_unit(const K2Ctx &cx, K3PWarp &ws, const int qn, const int part)
{
    constexpr int NALT = HALF ? 2 : 1;
    double acc[2][K3P_GP][2][NALT][2];
#pragma unroll
    for (int i = 0; i < 2 * K3P_GP * 2 * NALT * 2; ++i) (&acc[0][0][0][0][0])[i] = 0.0;
    k3p_mainloop<HALF>(acc, cx, ws, qn, K3P_GP * part);
    double vS[K3P_GP][4], vD[K3P_GP][4];
    const bool sw = K3T_SWAP && (ws.c & 2);
#pragma unroll
    for (int g = 0; g < K3P_GP; ++g)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double s0 = acc[0][g][0][0][h] + acc[1][g][0][0][h], s1 = acc[0][g][1][0][h] + acc[1][g][1][0][h];
            const double d0 = acc[0][g][0][NALT - 1][h] - acc[1][g][0][NALT - 1][h], d1 = acc[0][g][1][NALT - 1][h] - acc[1][g][1][NALT - 1][h];
            vS[g][2 * h] = sw ? s1 : s0; vS[g][2 * h + 1] = sw ? s0 : s1;
            vD[g][2 * h] = sw ? d1 : d0; vD[g][2 * h + 1] = sw ? d0 : d1;
        }
    k3p_emit<false>(ws, vS, qn, part);
    k3p_emit<true>(ws, vD, qn, part);
}

template <bool HALF>
__global__ void __launch_bounds__(K3P_THREADS, 1) k3_tmap_kernel(const K3PArgs a, const __grid_constant__ K3PMaps tm)
{
    extern __shared__ __align__(128) double sm[];
    const PlanDev &p = a.p;
    const int Q = p.Q, Kp = p.Kp, ldu = p.ldu, N = p.N;
    const K3PSmem L = k3p_smem_layout(Q, Kp, ldu, a.nslots);
    const int ldt = L.ldt, qp = L.qp;
    double *sUQ = sm + L.uq, *sMu = sm + L.mu, *sW = sm + L.w;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + L.bars);
    unsigned long long *trig_ready = bars, *trig_free = bars + 2, *uq_bar = bars + 4;
    volatile int *tok = reinterpret_cast<volatile int *>(sm + L.tok);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int upc = (((Q + 15) >> 4) + K3P_GP - 1) / K3P_GP;      // parts per column (1 or 2)

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
            for (int b = 0; b < 2; ++b) { mbar_init(trig_ready + b, K3P_WARPS); mbar_init(trig_free + b, K3P_WARPS); }
            mbar_init_fence();
            mbar_expect_tx(uq_bar, uq_bytes);
            bulk_load(sUQ, p.uq, uq_bytes, uq_bar);
        }
        if (tid < K3P_WARPS) tok[tid] = tid < 8 ? 1 : 0;           // two tokens per ring: with the warps 0..3 and 4..7
        for (int i = tid; i < Kp; i += K3P_THREADS) { sMu[i] = p.mu[i]; sW[i] = p.w[i]; }
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
    cx.bar_me = 0; cx.bar_other = 0;

    K3PWarp ws;
    ws.stg = sm + L.stg + warp * (a.nslots * K3P_SLOT);
    ws.ph = nullptr;
    ws.tm = &tm;
    const int nxt = (warp + 4) % K3P_WARPS;
    ws.tok_me = a.handoff ? tok + warp : nullptr;
    ws.tok_next = a.handoff ? tok + nxt : nullptr;
    ws.tok_init_next = nxt < 8 ? 1 : 0;
    ws.consumed = 0; ws.released = 0;
    ws.N = N; ws.Q = Q; ws.i0 = N - Q; ws.zskip = p.zskip; ws.qp = qp;
    ws.lane = lane; ws.r = lane >> 2; ws.c = lane & 3;
    ws.f0 = 0; ws.nslots = a.nslots; ws.cur = 0;
    ws.nostore = (a.flags & K3T_FLAG_NOSTORE) != 0;

    const double mhalf = 0.5 * (double)(p.twoj & 1);
    auto prefetch = [&](const long long kn) {
        const int bn = (int)(kn & 1);
        if (kn >= 2) mbar_wait(trig_free + bn, (unsigned)((kn - 2) >> 1) & 1u);
        const K2CItem itn = item_of(blockIdx.x + kn * G);
        double *tc = sm + L.trig + bn * L.trig_size, *ts = tc + K3T_ANG * ldt;
#pragma unroll 1
        for (int idx = tid; idx < K3T_ANG * Kp; idx += K3P_THREADS) {
            const int ang = idx / Kp, kp = idx - ang * Kp;
            long long b = itn.f0 + ang; if (b >= a.nb) b = a.nb - 1;
            double sn, cs;
            sincos(sMu[kp] * a.beta[b], &sn, &cs);
            const double w = sW[kp];
            tc[ang * ldt + kp] = w * cs;
            ts[ang * ldt + kp] = w * sn;
        }
        double *ph = sm + L.ph + bn * L.ph_size;
#pragma unroll 1
        for (int idx = tid; idx < 2 * K3T_ANG * Q; idx += K3P_THREADS) {
            const int which = idx / (K3T_ANG * Q), rem2 = idx - which * (K3T_ANG * Q);
            const int ang = rem2 / Q, q = rem2 - ang * Q;
            long long b = itn.f0 + ang; if (b >= a.nb) b = a.nb - 1;
            double sn, cs;
            sincos(((double)q + mhalf) * (which ? a.gamma[b] : a.alpha[b]), &sn, &cs);
            ph[((2 * which + 0) * K3T_ANG + ang) * qp + k3t_perm(q)] = cs;
            ph[((2 * which + 1) * K3T_ANG + ang) * qp + k3t_perm(q)] = sn;
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
        // units = (column, part) pairs in column order; the warps take adjacent units (rotated from item to item)
        const int v_begin = it.q_begin * upc, nunits = (it.q_end - it.q_begin) * upc;
        const int wslot = (warp + (int)(k % K3P_WARPS)) % K3P_WARPS;
        const int iters = (nunits + K3P_WARPS - 1) / K3P_WARPS;      // the warps of a ring make the same number of turns
        bool pre = k + 1 >= nmine;
#pragma unroll 1
        for (int u = 0; u < iters; ++u) {
            k3p_acquire(ws);
            const int v = u * K3P_WARPS + wslot;
            if (v < nunits) {
                const int vv = v_begin + v;
                k3p_unit<HALF>(cx, ws, vv / upc, vv - (vv / upc) * upc);
            } else k3p_release(ws);                                   // idle turn: pass the token on
            if (!pre) { prefetch(k + 1); pre = true; }
        }
        if (!pre) prefetch(k + 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(trig_free + b);
    }
    bulk_wait_all();
}

}  // namespace wd
