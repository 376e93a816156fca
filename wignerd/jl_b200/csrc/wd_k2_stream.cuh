// wd_k2_stream.cuh -- the DMMA contraction for j whose eigenvector table does not fit in shared memory
// (2j > 223, up to the supported maximum): same mathematics, warp unit, epilogue and store warps as
// k2_real_kernel (wd_k2_dmma.cuh), but the mu axis is streamed.
//
//   CTA tile = 16 angles x one block of 16*GMAX quadrant rows x 8 quadrant columns (one column per compute warp).
//   The mu axis is cut into chunks of <= 32 columns per eigenvector class; per chunk a stage of a 4-deep ring in
//   shared memory receives, through the TMA engine (cp.async.bulk -> SASS UBLKCP, completion on the stage's `full`
//   mbarrier; three copies per chunk, from a chunk-major copy of the table that has the stage's row layout),
//       - the row block of UQ      [16*GMAX rows][32 (+2 pad)]
//       - the 8 column rows of UQ  [8][32 (+2 pad)]
//       - the trig tile            [cos|sin][16 angles][32 (+4 pad)]   from a table precomputed once per call
//         (k2s_trig_kernel) -- in-kernel trig would be recomputed by every one of the ~Q^2/512 tiles of a chunk.
//   Producer/consumer pipeline without CTA barriers: chunk g is requested by compute warp g % 8 (after its work on
//   chunk g - 3), every compute warp waits on full[stage], and releases the stage by
//   arriving on empty[stage] (8 arrivals).  The first version -- every thread issuing ~7 cp.async per chunk with
//   their index arithmetic, cp.async.wait + a CTA named barrier per chunk -- spent 8 % of the compute warps' time in
//   the issue loop and 3 % at the barrier (ncu source view at j = 410).
//   Accumulators live in registers across the whole mu loop.
//   Shared-memory traffic per DMMA equals the resident kernel's; global->shared traffic is ~10 B/clk/SM.
#pragma once
#include "wd_k2_dmma.cuh"

namespace wd {

constexpr int K2S_KC = 32;                 // mu columns per chunk
constexpr int K2S_LDU = K2S_KC + 2;        // == 2 mod 4: conflict-free stride-2 row fragments
constexpr int K2S_LDT = K2S_KC + 4;        // == 4 mod 8: conflict-free 8-row fragments
constexpr int K2S_STAGES = 4;
constexpr int K2S_TRIG_TILE = 2 * K2_ANG * K2S_LDT;       // doubles per (angle chunk, mu chunk) trig tile

__host__ __device__ inline int k2s_nchunks(int Kxp) { return (Kxp + K2S_KC - 1) / K2S_KC; }

struct K2SArgs {
    K2Args base;
    const double *trig;        // [angle chunk][mu chunk][2][16][K2S_LDT]
    const double *uqc;         // chunk-major copy of UQ: [mu chunk][qpad rows][K2S_LDU], zero on the pads (k2s_table_kernel)
    int qpad;                  // rows per chunk of uqc: Q rounded up to whole row blocks
    int nck0, nck1;            // mu chunks of class 0 / class 1
    const int *pair_flag;      // device flag of k2c_pair_check_kernel (1: the batch is symmetric, beta[b] + beta[nb-1-b] = pi):
                               // only the first half of the angles is contracted, every tile also feeds the mirror matrices
};

// Chunk-major, row-padded copy of a plan's UQ: the row block and the column rows of a mu chunk are then contiguous
// in global memory in exactly the shared-memory stage layout, so that ONE bulk copy brings each of them in (72 copies
// of 256 bytes per chunk before: ~12 instructions each on the producing warp, ~9 % of the compute warps' samples).
__global__ void __launch_bounds__(256) k2s_table_kernel(PlanDev p, double *__restrict__ uqc, int qpad, int nck0, int nck1)
{
    const long long total = (long long)(nck0 + nck1) * p.Q * K2S_KC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % K2S_KC);
        const int q = (int)((i / K2S_KC) % p.Q);
        const int kc = (int)(i / ((long long)K2S_KC * p.Q));
        const int cls = kc >= nck0;
        const int kbase = cls ? p.K0p + (kc - nck0) * K2S_KC : kc * K2S_KC;
        const int kend = cls ? p.Kp : p.K0p;
        if (kbase + col < kend) uqc[((size_t)kc * qpad + q) * K2S_LDU + col] = p.uq[(size_t)q * p.ldu + kbase + col];
    }
}
__host__ __device__ inline int k2s_qpad(int Q, bool half) { const int rows = 16 * (half ? 2 : 4); return (Q + rows - 1) / rows * rows; }

struct K2SSmem {               // offsets in doubles
    int stage, stage_size, uqr, uqc, trig, stg, desc, bars, total;
    int rows;                  // 16 * GMAX
};
__host__ __device__ inline K2SSmem k2s_smem_layout(bool half)
{
    K2SSmem s;
    s.rows = 16 * (half ? 2 : 4);
    s.uqr = 0;
    s.uqc = s.rows * K2S_LDU;
    s.trig = s.uqc + K2_CWARPS * K2S_LDU;
    s.stage_size = s.trig + K2S_TRIG_TILE;
    s.stage_size = (s.stage_size + 1) & ~1;
    int o = 0;
    s.stage = o; o += K2S_STAGES * s.stage_size;
    s.stg = o;   o += K2_CWARPS * 2 * K2_HT_SIZE;                       // a 2-slot half-tile ring per compute warp
    s.desc = o;  o += K2_CWARPS * 2 * (int)(sizeof(K2Tile) / 8);
    s.bars = o;  o += 4 * K2_CWARPS + 2 * K2S_STAGES;                    // ring: full[8][2], empty[8][2]; stages: full[4], empty[4]
    s.total = o;
    return s;
}

// trig table: T[ac][kc][0|1][row(angle)][col] = w * cos|sin(mu * beta), zero on the pads
__global__ void __launch_bounds__(256) k2s_trig_kernel(PlanDev p, const double *__restrict__ beta, long long nb,
                                                       double *__restrict__ trig, int nck0, int nck1, const int *pair_flag)
{
    const int nck = nck0 + nck1;
    if (pair_flag != nullptr && *reinterpret_cast<const volatile int *>(pair_flag) != 0) nb = (nb + 1) / 2;   // front half only
    const long long nac = (nb + K2_ANG - 1) / K2_ANG;
    const long long total = nac * nck * K2_ANG * K2S_KC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % K2S_KC);
        const int ang = (int)((i / K2S_KC) % K2_ANG);
        const int kc = (int)((i / (K2S_KC * K2_ANG)) % nck);
        const long long ac = i / ((long long)K2S_KC * K2_ANG * nck);
        const int cls = kc >= nck0;
        const int kbase = cls ? p.K0p + (kc - nck0) * K2S_KC : kc * K2S_KC;
        const int kend = cls ? p.Kp : p.K0p;
        const int kp = kbase + col;
        long long b = ac * K2_ANG + ang; if (b >= nb) b = nb - 1;
        double c = 0.0, s = 0.0;
        if (kp < kend) {
            sincos(p.mu[kp] * beta[b], &s, &c);            // fl(mu*beta) then sincos, like src/WignerD.jl:199
            const double w = p.w[kp];
            c *= w; s *= w;
        }
        double *tile = trig + ((size_t)ac * nck + kc) * K2S_TRIG_TILE;
        const int row = k2_trig_row(ang);
        tile[row * K2S_LDT + col] = c;
        tile[(K2_ANG + row) * K2S_LDT + col] = s;
    }
}

// DMMAs of one mu chunk for one class: the two-fragment-set pipeline of k2_mainloop over nsteps k-steps.
template <bool HALF, int GMAX, int GC>
__device__ __forceinline__ void k2s_chunk(double (&acc)[2][GMAX][2][HALF ? 2 : 1][2], const double *tA, const double *tB,
                                          const double *un, const double *ub, const int nsteps)
{
    constexpr int gs = 16 * K2S_LDU, t8 = 8 * K2S_LDT, ldu = K2S_LDU;
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
        for (int ai = 0; ai < 2; ++ai) { f.ta[ai] *= f.un; f.tb[ai] *= f.un; }
#pragma unroll
        for (int g = 0; g < GC; ++g) {
#pragma unroll
            for (int ai = 0; ai < 2; ++ai) {
                dmma(acc[ai][g][0][0], f.ta[ai], f.be[g]);
                dmma(acc[ai][g][1][0], f.tb[ai], f.bo[g]);
                if (HALF) {
                    dmma(acc[ai][g][0][1], f.tb[ai], f.be[g]);
                    dmma(acc[ai][g][1][1], f.ta[ai], f.bo[g]);
                }
            }
        }
    };
    // the stage rows are K2S_KC wide and followed by pad columns inside the stage, so the look-ahead load of the
    // step after the last one stays inside the stage
    Frag f0, f1;
    load(f0, 0);
    int k = 0;
    const int kend = 4 * nsteps;
#pragma unroll 1
    for (; k + 4 < kend; k += 8) {
        load(f1, k + 4);
        step(f0);
        load(f0, (k + 8 < K2S_KC) ? k + 8 : 0);
        step(f1);
    }
    if (k < kend) step(f0);
}

template <bool HALF>
__global__ void __launch_bounds__(K2_THREADS_REAL, 1) k2_stream_kernel(K2SArgs sa)
{
    constexpr int GMAX = HALF ? 2 : 4;
    constexpr int NALT = HALF ? 2 : 1;
    constexpr int ROWS = 16 * GMAX;
    extern __shared__ __align__(128) double sm[];
    const K2Args &a = sa.base;
    const PlanDev &p = a.p;
    const int Q = p.Q, ldu = p.ldu;
    const K2SSmem L = k2s_smem_layout(HALF);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + L.bars);
    unsigned long long *sfull = bars + 4 * K2_CWARPS, *sempty = sfull + K2S_STAGES;
    if (tid == 0) {
        for (int w = 0; w < 4 * K2_CWARPS; ++w) mbar_init(bars + w, 1);
        for (int st = 0; st < K2S_STAGES; ++st) { mbar_init(sfull + st, 1); mbar_init(sempty + st, K2_CWARPS); }
        mbar_init_fence();
    }
    __syncthreads();
    if (warp >= K2_CWARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(K2_REGS_STORE));
        k2_store_loop(sm + L.stg, sm + L.desc, bars, warp - K2_CWARPS, a.out, (size_t)p.N * p.N, lane);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(K2_REGS_COMPUTE));

    K2Ctx cx;
    cx.sUQ = nullptr; cx.sTc = nullptr; cx.sTs = nullptr; cx.sPh = nullptr;
    cx.stg = nullptr;
    K2RCtx rc;
    rc.stg = sm + L.stg + warp * (2 * K2_HT_SIZE);
    rc.desc = reinterpret_cast<K2Tile *>(sm + L.desc) + 2 * warp;
    rc.full = bars + 2 * warp; rc.empty = bars + 2 * K2_CWARPS + 2 * warp;
    rc.tiles = 0;
    rc.NN = (size_t)p.N * p.N;
    const int paired = (sa.pair_flag != nullptr && a.nb >= 2) ? *reinterpret_cast<const volatile int *>(sa.pair_flag) : 0;
    const long long H = paired ? (a.nb + 1) / 2 : a.nb;          // angles that are contracted
    rc.paired = paired; rc.nrowsb0 = rc.nrowsb1 = 0; rc.e_item_b = 0;
    cx.ldu = ldu; cx.ldt = 0; cx.N = p.N; cx.Q = Q; cx.i0 = p.N - Q; cx.zskip = p.zskip;
    cx.Kp = p.Kp; cx.K0p = p.K0p; cx.twoj = p.twoj;
    cx.lane = lane; cx.r = lane >> 2; cx.c = lane & 3;
    cx.bar_me = 0; cx.bar_other = 0;
    const int r = cx.r, c = cx.c;

    const int ngroups = (Q + 15) >> 4;
    const int nrb = (ngroups + GMAX - 1) / GMAX;           // row blocks
    const int ncb = (Q + K2_CWARPS - 1) / K2_CWARPS;       // column blocks (8 columns each)
    const long long nac = (H + K2_ANG - 1) / K2_ANG;
    const long long ntiles = nac * nrb * ncb;
    const int nck0 = sa.nck0, nck = sa.nck0 + sa.nck1;
    unsigned gch = 0;                                      // chunks consumed so far by this CTA (ring position / phases)

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int cb = (int)(tile % ncb);
        const int rb = (int)((tile / ncb) % nrb);
        const long long ac = tile / ((long long)ncb * nrb);
        cx.b0 = ac * K2_ANG;
        rc.e_item = (size_t)cx.b0 * rc.NN;
        rc.nrows0 = (int)min(8LL, (H - cx.b0 + 1) / 2);
        rc.nrows1 = (int)min(8LL, (H - cx.b0) / 2);
        rc.ntiles = rc.nrows1 > 0 ? 2 : 1;
        if (paired) {
            // angle b0 + ai + 2 ar has the mirror matrix nb-1-b (if that lies in the second half): a prefix of the tile's angles
            const long long last = a.nb - 1 - H - cx.b0;          // largest angle offset whose mirror is >= H
            rc.nrowsb0 = last >= 0 ? (int)min((long long)rc.nrows0, last / 2 + 1) : 0;
            rc.nrowsb1 = last >= 1 ? (int)min((long long)rc.nrows1, (last - 1) / 2 + 1) : 0;
            rc.e_item_b = (size_t)(a.nb - 1 - cx.b0) * rc.NN;
        }
        const int g0 = rb * GMAX, gcnt = min(GMAX, ngroups - g0);
        const int qn = cb * K2_CWARPS + warp;
        const bool active = qn < Q;
        const int pe = qn & 1;

        // Producer: the whole warp loads mu chunk kc of this tile (global chunk number g) into stage g % STAGES.
        auto produce = [&](const int kc, const unsigned g) {
            const int si = (int)(g % K2S_STAGES);
            const unsigned use = g / K2S_STAGES;
            if (use > 0) mbar_wait(sempty + si, (use - 1) & 1u);         // all 8 warps are done with the stage's previous chunk
            double *st = sm + L.stage + si * L.stage_size;
            if (lane == 0) {
                // three bulk copies: the row block, the 8 column rows (both contiguous in the chunk-major table, pads
                // included: rows past Q are zero and never stored) and the trig tile
                constexpr unsigned row_bytes = K2S_LDU * 8u;
                mbar_expect_tx(sfull + si, (unsigned)(ROWS + K2_CWARPS) * row_bytes + (unsigned)K2S_TRIG_TILE * 8u);
                const double *src = sa.uqc + (size_t)kc * sa.qpad * K2S_LDU;
                bulk_load(st + L.uqr, src + (size_t)(16 * g0) * K2S_LDU, (unsigned)ROWS * row_bytes, sfull + si);
                bulk_load(st + L.uqc, src + (size_t)(cb * K2_CWARPS) * K2S_LDU, (unsigned)K2_CWARPS * row_bytes, sfull + si);
                bulk_load(st + L.trig, sa.trig + ((size_t)ac * nck + kc) * K2S_TRIG_TILE, (unsigned)K2S_TRIG_TILE * 8u, sfull + si);
            }
        };

        double acc[2][2][GMAX][2][NALT][2];
#pragma unroll
        for (int i = 0; i < 2 * 2 * GMAX * 2 * NALT * 2; ++i) (&acc[0][0][0][0][0][0])[i] = 0.0;

        // prologue: the first STAGES-1 chunks of the tile
        for (int i = 0; i < K2S_STAGES - 1 && i < nck; ++i)
            if (warp == (int)((gch + i) % K2_CWARPS)) produce(i, gch + i);
#pragma unroll 1
        for (int kc = 0; kc < nck; ++kc) {
            const unsigned g = gch + kc;
            const int si = (int)(g % K2S_STAGES);
            mbar_wait(sfull + si, (g / K2S_STAGES) & 1u);                // chunk kc has landed
            if (active) {
                const double *st = sm + L.stage + si * L.stage_size;
                const int cls = kc >= nck0;
                const int kbase = cls ? p.K0p + (kc - nck0) * K2S_KC : kc * K2S_KC;
                const int kend = cls ? p.Kp : p.K0p;
                const int nsteps = min(K2S_KC, kend - kbase) >> 2;
                const double *tc = st + L.trig + r * K2S_LDT + c, *ts = tc + K2_ANG * K2S_LDT;
                const double *tA = pe ? ts : tc, *tB = pe ? tc : ts;
                const double *un = st + L.uqc + warp * K2S_LDU + c;
                const double *ub = st + L.uqr + (2 * r) * K2S_LDU + c;
                // only the 16-row groups that exist in this row block issue DMMAs (the last block of a column is partial)
                auto run = [&](double (&ac)[2][GMAX][2][NALT][2]) {
                    if (gcnt == GMAX) k2s_chunk<HALF, GMAX, GMAX>(ac, tA, tB, un, ub, nsteps);
                    else if (gcnt == 1) k2s_chunk<HALF, GMAX, 1>(ac, tA, tB, un, ub, nsteps);
                    else if (gcnt == 2) k2s_chunk<HALF, GMAX, (GMAX >= 2 ? 2 : 1)>(ac, tA, tB, un, ub, nsteps);
                    else k2s_chunk<HALF, GMAX, (GMAX >= 3 ? 3 : 1)>(ac, tA, tB, un, ub, nsteps);
                };
                if (cls == 0) run(acc[0]); else run(acc[1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sempty + si);                      // this warp is done with the stage
            {
                // request chunk kc + STAGES-1 (into the stage of chunk kc-1) AFTER the work on chunk kc: by now every warp
                // has released that stage, the producer does not stall; the copy has two chunks of compute to land
                const int nx = kc + K2S_STAGES - 1;
                if (nx < nck && warp == (int)((gch + nx) % K2_CWARPS)) produce(nx, gch + nx);
            }
        }
        gch += (unsigned)nck;
        if (!active) continue;

        // ---- epilogue: the resident real-d kernel's (k2_emit: S and D staged once each, two mirror targets per tile) ----
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
    k2_send_marker(cx, rc, K2_TILE_END_ALL);                  // tell the store warp that this compute warp has finished
}

}  // namespace wd
