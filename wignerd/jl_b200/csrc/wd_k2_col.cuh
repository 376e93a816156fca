// wd_k2_col.cuh -- third-generation real-d contraction kernel (k2_col_kernel): same mathematics and DMMA roles as
// k2_real_kernel (wd_k2_dmma.cuh), re-organised around the one thing that limited it -- HOW the output reaches HBM.
//
// What the store microbenchmarks say (tools/store_pattern_mb2.cu, B200, cfg-2 = 10^5 matrices of 201 x 201):
//   * the round-1 drain (512-byte runs at 8-byte alignment, 8 matrices interleaved)                  3.7 TB/s
//   * whole 1608-byte columns, one warp per column                                                   4.0-4.7 TB/s
//   * whole columns where NO 32-byte sector is ever written by two different warps ("sector-owned")  5.6-5.8 TB/s
//   * the same through the TMA engine (cp.async.bulk.global.shared::cta), even from 4 warps per SM   5.8-6.0 TB/s
//   (a plain streaming write: 6.05 TB/s).  With N = 2j+1 odd every column starts 8 bytes further into a sector than
//   its neighbour; a sector shared by two columns that is completed by another warp at another time is what costs.
//
// Structure:
//   * work unit = ONE quadrant column qn x ALL quadrant rows x 8 angles (one MMA row tile): accumulators
//     [class 2][<= 7 groups of 16 rows][even/odd rows][2] = 56 doubles, the register budget of the old 16-angle x 64-row
//     unit.  A unit therefore owns two whole output columns per matrix: c0 = j+n ([D descending | S ascending]) and
//     c1 = j-n ([S descending | D ascending]).
//   * every warp computes AND stores (no store warps): the butterfly results are written, with their signs, into a
//     staging row per (half column, angle) that is laid out in GLOBAL sector phase (row index p <-> global element
//     floor4(es) + p), and leave through ONE bulk-TMA store per (half column, angle) that covers whole 32-byte sectors
//     only.  A warp owns a contiguous block of columns of an item and walks it in memory order, so the <= 3 elements of
//     the sector that a stretch shares with the NEXT stretch of the walk are held back in a small carry array and leave
//     with that stretch's store.  Only the two ends of a block are written with partial-sector stores.
//   * beta <-> pi - beta pairing (test/runtests.jl:330-334: d(pi-b)[m,n] = (-1)^(j+m) d(b)[m,-n]): when the
//     angle grid is symmetric (beta[b] + beta[nb-1-b] = pi for every b; checked on the device by
//     k2c_pair_check_kernel, no host round trip) one unit also writes the two columns of the mirror matrix nb-1-b
//     from the SAME accumulators: column c0 of the mirror matrix is column c1 of this one with alternating row signs.
//     The DMMA work per output element halves; the kernel becomes HBM-write bound.
//   * trig tables of the next item are made during the current one (as in k2_real_kernel); pair hand-off between the
//     two warps of an SM sub-partition as before.
//   Measured at cfg-2 (j = 100, 10^5 angles, one B200): symmetric grid 6.4 ms (k2_real_kernel: 9.1-9.25 ms) = 5.0 TB/s of
//   output; generic grid 9.5 ms -- the 8-angle unit re-reads the A operand twice as often and every warp pays for its own
//   stores, so generic grids stay with k2_real_kernel (9.1 ms): the host launches both kernels and the device-side
//   pair check decides which of the two does the work (K2CArgs::require_pair / K2Args::skip_flag).
#pragma once
#include "wd_k2_dmma.cuh"

namespace wd {

constexpr int K2C_ANG = 8;                  // angles per item = one MMA row tile
constexpr int K2C_WARPS = 8;
constexpr int K2C_THREADS = K2C_WARPS * 32;
constexpr int K2C_GMAX_INT = 7;             // 16-row groups per column: Q <= 112 (integer j)
constexpr int K2C_GMAX_HALF = 4;            // half-integer j carries the second trig function: Q <= 64
constexpr int K2C_FLAG_NOSTORE = 1;         // experiment bit: no global stores at all

struct K2CArgs {
    PlanDev p;
    const double *beta;
    long long nb;
    double *out;
    const int *pair_flag;     // device flag of k2c_pair_check_kernel (1: b <-> nb-1-b are mirror angles); nullptr = unpaired
    int handoff;              // see K2Args::handoff
    int flags;
    int require_pair;         // 1: do nothing unless the grid is paired (the unpaired batch goes to k2_real_kernel, launched alongside)
};

struct K2CSmem {                // offsets in doubles, identical on host and device
    int uq, mu, w, trig, stg, carry, bars, total;
    int ldt, rs, trig_size, slot;
};
__host__ __device__ inline K2CSmem k2c_smem_layout(int Q, int Kp, int ldu, int N)
{
    K2CSmem s;
    s.ldt = Kp | 4;
    // staging row of a stretch (half a column, <= Q elements): index p <-> global element floor4(es) + p; p < ph + Q + 3 <= Q + 6.
    // Odd N: consecutive angles have distinct phases and a stride == 0 (mod 4) keeps the staging stores conflict-free;
    // even N: stride == 2 (mod 4) (2-way conflicts: the TMA source must stay 16-byte aligned).
    s.rs = ((Q + 6) & ~3) + ((N & 1) ? 0 : 2);
    s.trig_size = 2 * K2C_ANG * s.ldt;
    s.slot = K2C_ANG * s.rs;
    int o = 0;
    s.uq = o;    o += Q * ldu;                       // no pad rows: row indices are clamped in the main loop
    s.mu = o;    o += Kp;
    s.w = o;     o += Kp;
    s.trig = o;  o += 2 * s.trig_size;
    o = (o + 3) & ~3;
    s.stg = o;   o += K2C_WARPS * 2 * s.slot;
    s.carry = o; o += K2C_WARPS * 4 * K2C_ANG * 4;   // [warp][target][angle][<= 3 held-back elements]
    s.bars = o;  o += 8;                             // trig_ready[2], trig_free[2], UQ load, (spare)
    s.total = o;
    return s;
}
__host__ __device__ inline bool k2c_shape_ok(int twoj)
{
    // N >= 6: every stretch (half a column, >= 3 elements) then reaches the sector boundaries on both sides of the sector it
    // starts in (the carry logic assumes it); smaller matrices take the plain kernel
    const int Q = (twoj + 2) / 2;
    return twoj >= 5 && ((Q + 15) >> 4) <= ((twoj & 1) ? K2C_GMAX_HALF : K2C_GMAX_INT);
}

// 1 iff every pair (b, nb-1-b) satisfies |beta[b] + beta[nb-1-b] - pi| <= tol (one block; NaNs fail the test)
__global__ void __launch_bounds__(1024) k2c_pair_check_kernel(const double *__restrict__ beta, long long nb, double tol, int *flag)
{
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    int mine = 0;
    const long long H = nb / 2;
    // (four independent pairs per thread and pass: the loop is a chain of global-load latencies, 30 us for 10^5 angles unrolled once)
#pragma unroll 4
    for (long long b = threadIdx.x; b < H; b += blockDim.x) {
        const double d = (beta[b] + beta[nb - 1 - b]) - 3.141592653589793;
        if (!(fabs(d) <= tol)) mine = 1;
    }
    if (nb & 1) { if (threadIdx.x == 0 && !(fabs(2.0 * beta[H] - 3.141592653589793) <= tol)) mine = 1; }
    if (mine) bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) *flag = bad ? 0 : 1;
}

__device__ __forceinline__ void bulk_store(const void *gdst, const void *ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct K2CWarp {                // per-warp state of the store path
    double *stg;                // two slots of K2C_ANG staging rows
    double *carry;              // [target 4][angle 8][4]
    double *out;
    unsigned long long ebase;   // out / 8
    unsigned long long NN;
    long long nb, H, f0;        // batch size, number of front angles (paired) or nb, first front angle of the item
    int na, nvB;                // front angles of the item (1..8); how many of them have a mirror matrix to write (paired)
    int paired;
    int N, Q, i0, zskip, rs, slot;
    int lane, r, c;
    int nvl;                    // valid rows (0..4) of this lane in the LAST 16-row group of a column
    unsigned steps;             // staging steps so far (slot = steps & 1)
    unsigned have;              // bit t: target t has a held-back sector (a stretch of this block has been staged)
    int nostore;
    int elem;                   // doubles per matrix element (1: real d).  A complex variant of this kernel (elem = 2, phase tables,
                                // 16-byte staging stores) was built and dropped: 5.0 ms at cfg-3 against 4.0 ms of k2_cplx_kernel -- every
                                // lane owns 4 consecutive rows, so its 16-byte stores hit 2 of 8 banks, and the phase multiplies
                                // (6 FP64 operations per element and target) share the pipe with the DMMAs of the partner warp
};

// global index (in doubles from address 0) of row 0 of column col of the matrix of angle x of the item
__device__ __forceinline__ unsigned long long k2c_e0(const K2CWarp &ws, const bool back, const int col, const int x)
{
    const long long b = back ? ws.nb - 1 - (ws.f0 + x) : ws.f0 + x;
    return ws.ebase + (unsigned long long)ws.elem * ((unsigned long long)b * ws.NN + (unsigned long long)(col * ws.N));
}

// Second half of a staging step, after the values of the stretch (rows row0 .. row0 + nrows - 1 of column col) have been
// written to the slot: splice the held-back sector of the previous stretch of the target's chain in, hold this stretch's
// last partial sector back, and send the whole sectors off with one bulk-TMA store per angle.
__device__ __forceinline__ void k2c_splice_issue(K2CWarp &ws, double *slot, const bool back, const int col, const int row0,
                                                 const int nrows, const int t, const bool chain_asc, const bool have)
{
    const int lane = ws.lane, len = ws.elem * nrows, off0 = ws.elem * row0;
    __syncwarp();
    const int nv = back ? ws.nvB : ws.na;                             // angles of the item that exist for this target
    if (lane < 3 * K2C_ANG) {                                         // lane -> (angle lane / 3, element lane % 3)
        const int x = lane / 3, e = lane - 3 * x;
        if (x < nv) {
            const unsigned long long es = k2c_e0(ws, back, col, x) + off0;
            const int ph = (int)(es & 3ull), pe = ph + len;
            double *row = slot + x * ws.rs;
            double *cy = ws.carry + (t * K2C_ANG + x) * 4 + e;
            double *g0 = ws.out + (es - ws.ebase);                    // first double of the stretch in global memory
            if (chain_asc) {
                // carry-in: the ph doubles in front of the stretch; carry-out: its last pe & 3 doubles
                if (e < ph) { if (have) row[e] = *cy; }
                if (!have && ph > 0 && ph + e < 4 && !ws.nostore) g0[e] = row[ph + e];          // first stretch of the block: partial head sector
                if (e < (pe & 3)) *cy = row[(pe & ~3) + e];
            } else {
                const int nin = (4 - (pe & 3)) & 3;
                if (e < nin) { if (have) row[pe + e] = *cy; }
                if (!have && e < (pe & 3) && !ws.nostore) g0[len - (pe & 3) + e] = row[(pe & ~3) + e];   // first stretch: partial tail sector
                if (ph > 0 && ph + e < 4) *cy = row[ph + e];
            }
        }
    }
    fence_async_smem();
    __syncwarp();
    // lane -> angle.  (The compiler serialises the 8 copies -- UBLKCP takes uniform registers: ~11 instructions per copy;
    // tried and dropped, against 6.4-6.5 ms at cfg-2: one lane issuing all copies with shuffled parameters 7.4 ms; plain
    // 16-byte vector stores of the staged windows instead of the TMA 17.9 ms; a helper warp per compute warp (208 / 48
    // registers, full/empty mbarriers per slot) that splices the carries and issues the copies so that the compute warp only
    // writes its values: 7.5 ms, and 6.2 instead of 4.9 ms with the stores switched off -- eight more warps waiting on
    // mbarriers cost the DMMA warps more issue slots than the issue loop they took over.)
    if (lane < nv && !ws.nostore) {
        const unsigned long long es = k2c_e0(ws, back, col, lane) + off0;
        const int ph = (int)(es & 3ull), pe = ph + len;
        int s0, s1;
        if (chain_asc) { s0 = (ph > 0 && !have) ? 4 : 0; s1 = pe & ~3; }
        else           { s0 = ph > 0 ? 4 : 0;            s1 = have ? (pe + 3) & ~3 : pe & ~3; }
        if (s1 > s0) bulk_store(ws.out + ((es - ph) - ws.ebase) + s0, slot + lane * ws.rs + s0, (unsigned)(s1 - s0) * 8u);
    }
    bulk_commit();
    ws.steps++;
    ws.have |= 1u << t;
}
// One staging step = one STRETCH (half a column: its ascending part, rows i0 .. N-1 <- quadrant rows q = 0 .. Q-1, or its
// descending part, rows i0-1 .. 0 <- q = zskip .. Q-1) of one target column, for the 8 angles of the item.  Every lane
// writes its 4 consecutive quadrant rows per 16-row group, sign applied, into the staging row of its angle (laid out in
// global sector phase); 24 lanes then splice the held-back sector of the previous stretch of the target's chain in and
// hold this stretch's last partial sector back; 8 lanes issue one bulk-TMA store each.  Stretches of a target follow each
// other in memory order (chain_asc: ascending addresses), so only whole sectors are ever stored -- except at the two ends
// of the warp's column block.
//   v: the butterfly result that feeds this part; colsig: the column whose sigma applies (the front matrix's column).
struct K2CVal {                 // a butterfly result held as (lo, hi) words: the sign of a target is XORed into a copy of hi
    unsigned lo, hi;
};

template <bool DESCROWS, int GC>
__device__ __forceinline__ void k2c_stage(K2CWarp &ws, const K2CVal (&v)[GC][4], const bool back,
                                          const int col, const int colsig, const int t, const bool chain_asc)
{
    const int Q = ws.Q, i0 = ws.i0;
    const int row0 = DESCROWS ? 0 : i0, len = DESCROWS ? i0 : Q;
    // sign masks of the lane's rows jj = 0..3 of a group (16 g + 4 c vanish mod 4): sigma(row - colsig) [* (-1)^row for the
    // mirror matrix].  (Flipping the values in place, relative to their previous sign, saves a move per element but costs
    // 30 registers and spills: 7.9 instead of 6.5 ms at cfg-2.)
    unsigned m[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int row = DESCROWS ? Q - 1 - jj : i0 + jj;
        m[jj] = sigma_bit(row - colsig) ^ (back ? ((unsigned)row & 1u) << 31 : 0u);
    }
    const bool have = (ws.have >> t) & 1u;
    double *slot = ws.stg + (ws.steps & 1u) * ws.slot;
    bulk_wait_read<1>();                                              // the stores that read this slot two steps ago are done
    __syncwarp();
    {
        const int ph = (int)((k2c_e0(ws, back, col, ws.r) + row0) & 3ull);
        double *row = slot + ws.r * ws.rs + ph;
        double *pp = DESCROWS ? row + (Q - 1) - 4 * ws.c : row + 4 * ws.c;
        const bool skip0 = DESCROWS && ws.zskip && ws.c == 0;           // q = 0 is m = 0: it mirrors onto itself
#pragma unroll
        for (int g = 0; g < GC; ++g) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const bool ok = (g < GC - 1 || jj < ws.nvl) && !(g == 0 && jj == 0 && skip0);      // only the last group is partial
                if (ok) pp[DESCROWS ? -(16 * g + jj) : 16 * g + jj] = __hiloint2double((int)(v[g][jj].hi ^ m[jj]), (int)v[g][jj].lo);
            }
        }
    }
    k2c_splice_issue(ws, slot, back, col, row0, len, t, chain_asc, have);
}

// end of a warp's column block: the held-back sectors leave through partial stores.  q_last: last quadrant column of the block.
__device__ __forceinline__ void k2c_flush(K2CWarp &ws, const int q_last)
{
    __syncwarp();
    const int t = ws.lane >> 3, x = ws.lane & 7;
    const bool back = t >= 2, asc = (t == 0 || t == 2);      // targets: 0 front/c0, 1 front/c1, 2 back/c0, 3 back/c1; c0 chains ascend
    if (((ws.have >> t) & 1u) && x < (back ? ws.nvB : ws.na) && !ws.nostore) {
        const int col = asc ? ws.i0 + q_last : ws.Q - 1 - q_last;
        // last stretch of the chain: ascending chains end with the ascending part of the column, descending ones with the descending part
        const int row0 = ws.elem * (asc ? ws.i0 : 0), len = ws.elem * (asc ? ws.Q : ws.i0);
        const unsigned long long es = k2c_e0(ws, back, col, x) + row0;
        const int ph = (int)(es & 3ull), pe = ph + len;
        double *g0 = ws.out + (es - ws.ebase);
        const double *cy = ws.carry + (t * K2C_ANG + x) * 4;
        if (asc) { const int n = pe & 3; for (int e = 0; e < n; ++e) g0[len - n + e] = cy[e]; }
        else     { const int n = (4 - ph) & 3; for (int e = 0; e < n; ++e) g0[e] = cy[e]; }
    }
    ws.have = 0;
    __syncwarp();
}

// main loop of one unit: quadrant column qn, GC groups of 16 rows (all rows of the column), 8 angles
template <bool HALF, int GC>
__device__ __forceinline__ void k2c_mainloop(double (&acc)[2][GC][2][HALF ? 2 : 1][2], const K2Ctx &cx, const int qn)
{
    const int ldu = cx.ldu, ldt = cx.ldt, r = cx.r, c = cx.c, Q = cx.Q;
    const int pe = qn & 1;
    const double *tA = (pe ? cx.sTs : cx.sTc) + r * ldt + c;      // trig for the even-row tiles
    const double *tB = (pe ? cx.sTc : cx.sTs) + r * ldt + c;      // trig for the odd-row tiles
    const double *un = cx.sUQ + qn * ldu + c;
    // rows 16 g + 2 r (even tile) and + 1 (odd tile); only the last group can reach past the last row: its two row indices
    // are clamped (rows >= Q are computed on duplicate data and never stored)
    const double *ub = cx.sUQ + (2 * r) * ldu + c;
    const int gs = 16 * ldu;
    const double *ube = cx.sUQ + min(16 * (GC - 1) + 2 * r, Q - 1) * ldu + c;
    const double *ubo = cx.sUQ + min(16 * (GC - 1) + 2 * r + 1, Q - 1) * ldu + c;
#pragma unroll
    for (int cls = 0; cls < 2; ++cls) {
        const int kbeg = cls ? cx.K0p : 0;
        const int kend = cls ? cx.Kp : cx.K0p;
        struct Frag { double un, ta, tb, be[GC], bo[GC]; };
        auto load = [&](Frag &f, const int k) {
            f.un = un[k]; f.ta = tA[k]; f.tb = tB[k];
#pragma unroll
            for (int g = 0; g < GC - 1; ++g) { f.be[g] = ub[g * gs + k]; f.bo[g] = ub[g * gs + ldu + k]; }
            f.be[GC - 1] = ube[k]; f.bo[GC - 1] = ubo[k];
        };
        auto step = [&](Frag &f) {
            f.ta *= f.un; f.tb *= f.un;                               // A = T(beta, kp) * UQ[qn][kp]
#pragma unroll
            for (int g = 0; g < GC; ++g) {
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
        const int krel = !cx.bar_other ? -1 : (cls == 0 ? (cx.krel < 100 ? kbeg + 8 * cx.krel : -1)
                                                           : (cx.krel >= 100 ? kbeg + 8 * (cx.krel - 100) : -1));
        bool released = false;
#pragma unroll 1
        for (; k + 4 < kend; k += 8) {
            if (k == krel) { named_bar_arrive(cx.bar_other, 64); released = true; }
            load(f1, k + 4);
            step(f0);
            load(f0, k + 8);          // past the last step: reads (and discards) in-bounds shared memory
            step(f1);
        }
        if (k < kend) step(f0);
        if (cx.bar_other && !released && ((cls == 0 && cx.krel < 100) || (cls == 1 && cx.krel >= 100))) named_bar_arrive(cx.bar_other, 64);
    }
}

template <bool HALF, int GC>
__device__ __forceinline__ void k2c_unit(const K2Ctx &cx, K2CWarp &ws, const int qn)
{
    constexpr int NALT = HALF ? 2 : 1;
    double acc[2][GC][2][NALT][2];
#pragma unroll
    for (int i = 0; i < 2 * GC * 2 * NALT * 2; ++i) (&acc[0][0][0][0][0])[i] = 0.0;
    k2c_mainloop<HALF, GC>(acc, cx, qn);
    // butterfly: S = P0 + P1 -> (i,i'), (ir,i'r);  D = P0 - P1 -> (ir,i'), (i,i'r)  (half-integer j: D from the other trig function)
    K2CVal vS[GC][4], vD[GC][4];                        // [group][jj]: the lane's rows 16 g + 4 c + jj, jj = 2 h + eo
#pragma unroll
    for (int g = 0; g < GC; ++g)
#pragma unroll
        for (int eo = 0; eo < 2; ++eo)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double sv = acc[0][g][eo][0][h] + acc[1][g][eo][0][h];
                const double dv = acc[0][g][eo][NALT - 1][h] - acc[1][g][eo][NALT - 1][h];
                vS[g][2 * h + eo].lo = (unsigned)__double2loint(sv); vS[g][2 * h + eo].hi = (unsigned)__double2hiint(sv);
                vD[g][2 * h + eo].lo = (unsigned)__double2loint(dv); vD[g][2 * h + eo].hi = (unsigned)__double2hiint(dv);
            }
    const int c0 = ws.i0 + qn, c1 = ws.Q - 1 - qn;
    const bool mirror = qn >= ws.zskip;                     // n = 0 mirrors onto itself: no column c1
    // front matrix b: column c0 = [D descending | S ascending], column c1 = [S descending | D ascending];
    // mirror matrix nb-1-b: its column c0 is the front matrix's column c1 with alternating row signs, its c1 the front's c0.
    // The chains of c0 ascend in memory (descending part first), those of c1 descend (ascending part first).
    k2c_stage<true, GC>(ws, vD, false, c0, c0, 0, true);
    k2c_stage<false, GC>(ws, vS, false, c0, c0, 0, true);
    if (mirror) {
        k2c_stage<false, GC>(ws, vD, false, c1, c1, 1, false);
        k2c_stage<true, GC>(ws, vS, false, c1, c1, 1, false);
    }
    if (ws.paired) {
        k2c_stage<true, GC>(ws, vS, true, c0, c1, 2, true);
        k2c_stage<false, GC>(ws, vD, true, c0, c1, 2, true);
        if (mirror) {
            k2c_stage<false, GC>(ws, vS, true, c1, c0, 3, false);
            k2c_stage<true, GC>(ws, vD, true, c1, c0, 3, false);
        }
    }
}

struct K2CItem { long long f0; int na, q_begin, q_end; };

template <bool HALF, int GC>
__global__ void __launch_bounds__(K2C_THREADS, 1) k2_col_kernel(K2CArgs a)
{
    extern __shared__ __align__(128) double sm[];
    const PlanDev &p = a.p;
    const int Q = p.Q, Kp = p.Kp, ldu = p.ldu, N = p.N;
    const K2CSmem L = k2c_smem_layout(Q, Kp, ldu, N);
    const int ldt = L.ldt;
    double *sUQ = sm + L.uq, *sMu = sm + L.mu, *sW = sm + L.w;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + L.bars);
    unsigned long long *trig_ready = bars, *trig_free = bars + 2, *uq_bar = bars + 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // items: chunks of 8 (front) angles; paired: only the first half of the batch, every unit also writes the mirror matrices
    const int paired = (a.pair_flag != nullptr && a.nb >= 2) ? *reinterpret_cast<const volatile int *>(a.pair_flag) : 0;
    if (a.require_pair && !paired) return;
    const long long H = paired ? (a.nb + 1) / 2 : a.nb;
    const long long nch = (H + K2C_ANG - 1) / K2C_ANG, G = gridDim.x;
    const long long nfull = (nch / G) * G, rem = nch - nfull;
    const int tsplit = rem ? (int)max(1LL, min(16LL, G / rem)) : 1;       // the last, partial wave is split over the idle CTAs
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
        o.f0 = ch * K2C_ANG;
        o.na = (int)min((long long)K2C_ANG, H - o.f0);
        return o;
    };

    {
        const unsigned uq_bytes = (unsigned)(Q * ldu) * 8u;        // a multiple of 16: ldu is even
        if (tid == 0) {
            mbar_init(uq_bar, 1);
            for (int b = 0; b < 2; ++b) { mbar_init(trig_ready + b, K2C_WARPS); mbar_init(trig_free + b, K2C_WARPS); }
            mbar_init_fence();
            mbar_expect_tx(uq_bar, uq_bytes);
            bulk_load(sUQ, p.uq, uq_bytes, uq_bar);
        }
        for (int i = tid; i < Kp; i += K2C_THREADS) { sMu[i] = p.mu[i]; sW[i] = p.w[i]; }
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
    cx.bar_me = a.handoff ? 2 + warp : 0;
    cx.bar_other = a.handoff ? 2 + (warp ^ 4) : 0;
    if (a.handoff && (warp & 4)) named_bar_arrive(cx.bar_other, 64);       // warps 0..3 take the first turn

    K2CWarp ws;
    ws.stg = sm + L.stg + warp * (2 * L.slot);
    ws.carry = sm + L.carry + warp * (4 * K2C_ANG * 4);
    ws.out = a.out;
    ws.ebase = (unsigned long long)(uintptr_t)a.out >> 3;
    ws.NN = (unsigned long long)N * N;
    ws.nb = a.nb; ws.H = H; ws.paired = paired;
    ws.N = N; ws.Q = Q; ws.i0 = N - Q; ws.zskip = p.zskip; ws.rs = L.rs; ws.slot = L.slot;
    ws.lane = lane; ws.r = lane >> 2; ws.c = lane & 3;
    ws.nvl = max(0, min(4, Q - 16 * (((Q + 15) >> 4) - 1) - 4 * (lane & 3)));
    ws.steps = 0; ws.have = 0;
    ws.nostore = (a.flags & K2C_FLAG_NOSTORE) != 0;
    ws.elem = 1;

    // trig tables of item k+1 are made during item k (k2_real_kernel's scheme): Tc/Ts[angle][kp] = w cos|sin(mu beta),
    // fl(mu*beta) then sincos like src/WignerD.jl:199
    auto prefetch = [&](const long long kn) {
        const int bn = (int)(kn & 1);
        if (kn >= 2) mbar_wait(trig_free + bn, (unsigned)((kn - 2) >> 1) & 1u);
        const K2CItem itn = item_of(blockIdx.x + kn * G);
        double *tc = sm + L.trig + bn * L.trig_size, *ts = tc + K2C_ANG * ldt;
#pragma unroll 1
        for (int idx = tid; idx < K2C_ANG * Kp; idx += K2C_THREADS) {
            const int ang = idx / Kp, kp = idx - ang * Kp;
            long long b = itn.f0 + ang; if (b >= a.nb) b = a.nb - 1;
            double sn, cs;
            sincos(sMu[kp] * a.beta[b], &sn, &cs);
            const double w = sW[kp];
            tc[ang * ldt + kp] = w * cs;
            ts[ang * ldt + kp] = w * sn;
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
        cx.sTc = sm + L.trig + b * L.trig_size; cx.sTs = cx.sTc + K2C_ANG * ldt;
        ws.f0 = it.f0; ws.na = it.na;
        ws.nvB = paired ? (int)max(0LL, min((long long)it.na, a.nb - it.f0 - H)) : 0;
        // this warp's block of columns (rotated from item to item so that the uneven block sizes even out)
        const int ncols = it.q_end - it.q_begin, wslot = (warp + (int)(k & 7)) & 7;
        const int q_lo = it.q_begin + ncols * wslot / K2C_WARPS, q_hi = it.q_begin + ncols * (wslot + 1) / K2C_WARPS;
        const int iters = (ncols + K2C_WARPS - 1) / K2C_WARPS;        // the two warps of a pair make the same number of turns
        bool pre = k + 1 >= nmine;
#pragma unroll 1
        for (int u = 0; u < iters; ++u) {
            if (cx.bar_me) named_bar_sync(cx.bar_me, 64);
            const int qn = q_lo + u;
            if (qn < q_hi) k2c_unit<HALF, GC>(cx, ws, qn);
            else if (cx.bar_other) named_bar_arrive(cx.bar_other, 64);        // idle turn: pass the token on
            if (!pre) { prefetch(k + 1); pre = true; }
        }
        if (!pre) prefetch(k + 1);
        if (q_hi > q_lo) k2c_flush(ws, q_hi - 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(trig_free + b);
    }
    bulk_wait_all();
}

}  // namespace wd
