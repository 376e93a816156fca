// wd_api.cu -- host dispatcher + C ABI of libwignerd_b200.so (see include/wignerd_b200.h).
//
// Layering: C ABI -> handle (device, stream, per-j plan cache, mutex) -> kernels in wd_kernels.cuh.
// No CPU compute path exists in this file: every entry point that produces d/D values launches
// sm_100a kernels, and fails with WD_ERR_CUDA when no device is usable.
#include "../../../include/wignerd_b200.h"
#include "wd_kernels.cuh"
#include "wd_k2_dmma.cuh"
#include "wd_k2_stream.cuh"
#include "wd_k2_col.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace wd;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? WD_ERR_NOMEM : WD_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                  \
    } while (0)

constexpr int kMaxTwoJ = 4094;

struct Plan {
    PlanDev dev{};
    double *uq = nullptr, *mu = nullptr, *w = nullptr;
    double *uqc = nullptr;          // chunk-major copy of uq for the streaming kernel, built on first use
    int uqc_qpad = 0;
};

}  // namespace

struct wd_handle_s {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    std::mutex mtx;
    std::map<int, Plan> plans;
    int num_sms = 0;
    size_t smem_optin = 0;
    int64_t launches = 0;
    int variant = 0;
    int pairing = 1;                 // beta <-> pi - beta pairing in k2_col_kernel (wd_set_option)
    PlanDev *d_table = nullptr;      // plan table indexed by twoj, for K4
    int table_max = -1;
    bool table_dirty = true;
    int *d_flag = nullptr;           // error flag / scratch ints
    int *d_present = nullptr;        // [kMaxTwoJ + 2] j-presence flags of a tuple list (jmn path)
    bool attr_set_cplx[2] = {false, false};
    bool attr_set_stream[2] = {false, false};
    // stream-ordered pool for the streaming kernel's per-call trig tables, release threshold "never": with the device's
    // default pool the memory goes back to the OS at every synchronisation and every call after one pays for a fresh
    // mapping (0.45 -> 0.27 ms per call at j = 150 on 512 angles).  nullptr: creation failed, the default pool is used.
    cudaMemPool_t pool = nullptr;
    bool attr_set_real[2] = {false, false};
    int *d_pair = nullptr;            // ring of device flags written by k2c_pair_check_kernel (one per call in flight)
    unsigned pair_idx = 0;
    // host-path workspace, kept across calls (grow-only): angle vectors and two output slabs + their events
    double *ws_ang = nullptr, *ws_slab[2] = {nullptr, nullptr};
    size_t ws_ang_bytes = 0, ws_slab_bytes[2] = {0, 0};
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
};

namespace {

void plan_dims(int twoj, PlanDev &p)
{
    p.twoj = twoj;
    p.N = twoj + 1;
    p.Q = (p.N + 1) / 2;
    const int K = p.Q;
    p.K0 = (K + 1) / 2;
    p.K1 = K / 2;
    p.K0p = wd_round4(p.K0);
    p.K1p = wd_round4(p.K1);
    p.Kp = p.K0p + p.K1p;
    p.ldu = p.Kp + 2;
    p.zskip = (twoj & 1) ? 0 : 1;
}

// One launch of the batched eigensolver for the (missing) twoj values todo[0..n).  Temporaries and half-built plans are
// released on every error path.
int prepare_batch(wd_handle_t h, const int *todo, size_t n)
{
    std::vector<EigJob> jobs(n);
    std::vector<Plan> fresh(n);
    size_t scratch_total = 0;
    int maxK = 1;
    for (size_t t = 0; t < n; ++t) {
        plan_dims(todo[t], fresh[t].dev);
        scratch_total += 2 * (size_t)fresh[t].dev.N * fresh[t].dev.Q;
        maxK = std::max(maxK, fresh[t].dev.Q);
    }
    double *scratch = nullptr;
    EigJob *d_jobs = nullptr;
    auto cleanup = [&](bool drop_tables) {
        if (scratch) cudaFree(scratch);
        if (d_jobs) cudaFree(d_jobs);
        if (drop_tables) for (auto &pl : fresh) if (pl.uq) cudaFree(pl.uq);
    };
    cudaError_t e = cudaMalloc(&scratch, scratch_total * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_jobs, n * sizeof(EigJob));
    if (e != cudaSuccess) { cleanup(true); return fail(WD_ERR_NOMEM, "cudaMalloc of the eigensolver workspace (%zu bytes) failed: %s", scratch_total * sizeof(double), cudaGetErrorString(e)); }
    size_t off = 0;
    for (size_t t = 0; t < n; ++t) {
        Plan &pl = fresh[t];
        const PlanDev &d = pl.dev;
        const size_t nuq = (size_t)d.Q * d.ldu;
        double *tables = nullptr;                       // one allocation per plan keeps wd_cache_clear simple
        e = cudaMalloc(&tables, (nuq + 2 * (size_t)d.Kp) * sizeof(double));
        if (e != cudaSuccess) { cleanup(true); return fail(WD_ERR_NOMEM, "cudaMalloc of the tables of 2j = %d failed: %s", d.twoj, cudaGetErrorString(e)); }
        pl.uq = tables;
        pl.mu = tables + nuq;
        pl.w = pl.mu + d.Kp;
        pl.dev.uq = pl.uq; pl.dev.mu = pl.mu; pl.dev.w = pl.w;
        cudaMemsetAsync(pl.uq, 0, nuq * sizeof(double), h->stream);
        EigJob &jb = jobs[t];
        jb.twoj = d.twoj; jb.N = d.N; jb.K = d.Q; jb.K0p = d.K0p; jb.ldu = d.ldu; jb.Kp = d.Kp;
        jb.uq = pl.uq; jb.mu = pl.mu; jb.w = pl.w;
        jb.scratch = scratch + off;
        off += 2 * (size_t)d.N * d.Q;
    }
    e = cudaMemcpyAsync(d_jobs, jobs.data(), n * sizeof(EigJob), cudaMemcpyHostToDevice, h->stream);
    for (size_t j0 = 0; j0 < n && e == cudaSuccess; j0 += 32768) {      // grid.y is limited to 65535: launch in slices
        const unsigned ny = (unsigned)std::min<size_t>(32768, n - j0);
        dim3 grid((maxK + 127) / 128, ny);
        k1_eig_kernel<<<grid, 128, 0, h->stream>>>(d_jobs + j0);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { cleanup(true); return fail(WD_ERR_CUDA, "batched eigensolver failed: %s", cudaGetErrorString(e)); }
    cleanup(false);
    for (size_t t = 0; t < n; ++t) h->plans[todo[t]] = fresh[t];
    h->table_dirty = true;
    return WD_OK;
}

// Build the tables for all missing twoj of the list; one launch of the batched eigensolver per <= 1 GiB of scratch.
int prepare_locked(wd_handle_t h, const int32_t *list, int nj)
{
    std::vector<int> todo;
    std::vector<char> seen((size_t)kMaxTwoJ + 1, 0);
    for (int i = 0; i < nj; ++i) {
        const int tj = list[i];
        if (tj < 0) return fail(WD_ERR_ASSERT, "j must be >= 0 (got 2j = %d)", tj);
        if (tj > kMaxTwoJ) return fail(WD_ERR_UNSUPPORTED, "2j = %d exceeds the supported maximum %d", tj, kMaxTwoJ);
        if (!seen[tj] && !h->plans.count(tj)) { seen[tj] = 1; todo.push_back(tj); }
    }
    if (todo.empty()) return WD_OK;
    CU(cudaSetDevice(h->device));
    const size_t cap = (size_t)1 << 27;                     // doubles of scratch per launch (1 GiB)
    size_t i = 0;
    while (i < todo.size()) {
        size_t j = i, scratch = 0;
        while (j < todo.size()) {
            const size_t need = (size_t)(todo[j] + 1) * (size_t)(todo[j] + 2);     // 2 N Q
            if (j > i && scratch + need > cap) break;
            scratch += need; ++j;
        }
        const int rc = prepare_batch(h, todo.data() + i, j - i);
        if (rc) return rc;
        i = j;
    }
    return WD_OK;
}

int get_plan(wd_handle_t h, int twoj, const Plan **out)
{
    int32_t tj = twoj;
    int rc = prepare_locked(h, &tj, 1);
    if (rc) return rc;
    *out = &h->plans[twoj];
    return WD_OK;
}

// Resident DMMA kernel for complex D (k2_cplx_kernel)
bool cplx_fits(wd_handle_t h, const PlanDev &p, size_t *bytes)
{
    const K2Smem L = k2_smem_layout(p.Q, p.Kp, p.ldu);
    *bytes = (size_t)L.total * sizeof(double);
    return *bytes <= h->smem_optin;
}

template <bool HALF>
int launch_cplx(wd_handle_t h, const K2Args &a, size_t smem)
{
    if (!h->attr_set_cplx[HALF]) {
        CU(cudaFuncSetAttribute(k2_cplx_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
        h->attr_set_cplx[HALF] = true;
    }
    const long long nchunks = (a.nb + K2_ANG - 1) / K2_ANG;
    // fewer chunks than SMs: split every chunk's units over several CTAs (each rebuilds the chunk's small trig tables)
    K2Args b = a;
    b.nsplit = nchunks >= h->num_sms ? 1 : (int)std::min<long long>(16, (h->num_sms + nchunks - 1) / nchunks);
    const int grid = (int)std::min<long long>(nchunks * b.nsplit, h->num_sms);
    k2_cplx_kernel<HALF><<<grid, K2_THREADS_CPLX, smem, h->stream>>>(b);
    h->launches++;
    CU(cudaGetLastError());
    return WD_OK;
}

// Resident DMMA kernel for real d (k2_real_kernel)
bool real_fits(wd_handle_t h, const PlanDev &p, size_t *bytes)
{
    const K2RSmem L = k2r_smem_layout(p.Q, p.Kp, p.ldu);
    *bytes = (size_t)L.total * sizeof(double);
    return *bytes <= h->smem_optin;
}

template <bool HALF>
int launch_real(wd_handle_t h, const K2Args &a, size_t smem)
{
    if (!h->attr_set_real[HALF]) {
        CU(cudaFuncSetAttribute(k2_real_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
        h->attr_set_real[HALF] = true;
    }
    // items: whole chunks of 16 angles for every full wave of the grid; the chunks of the last, partial wave (or of a
    // batch smaller than the GPU) are split over several CTAs each (every CTA rebuilds the chunk's small trig tables)
    K2Args b = a;
    const long long G = h->num_sms;
    const long long nchunks = (a.nb + K2_ANG - 1) / K2_ANG;
    b.nfull = (nchunks / G) * G;
    const long long rem = nchunks - b.nfull;
    b.tsplit = rem ? (int)std::max<long long>(1, std::min<long long>(16, G / rem)) : 1;
    {
        const char *e = getenv("WD_K2_TAIL");             // kernel-experiment knob: 0 = never split the last wave
        if (e && atoi(e) == 0 && b.nfull > 0) b.tsplit = 1;
    }
    b.nitems = b.nfull + rem * b.tsplit;
    const int grid = (int)std::min<long long>(b.nitems, G);
    k2_real_kernel<HALF><<<grid, K2_THREADS_REAL, smem, h->stream>>>(b);
    h->launches++;
    CU(cudaGetLastError());
    return WD_OK;
}

// Column-staged DMMA kernel for real d (k2_col_kernel, wd_k2_col.cuh): the default for tables that fit in shared memory
constexpr int kPairRing = 64;
bool col_fits(wd_handle_t h, const PlanDev &p, size_t *bytes)
{
    const K2CSmem L = k2c_smem_layout(p.Q, p.Kp, p.ldu, p.N);
    *bytes = (size_t)L.total * sizeof(double);
    return k2c_shape_ok(p.twoj) && *bytes <= h->smem_optin;
}

template <bool HALF, int GC>
cudaError_t launch_col_gc(const K2CArgs &c, int grid, size_t smem, size_t optin, cudaStream_t st)
{
    static thread_local bool attr_set = false;            // per instantiation; setting it again is harmless
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k2_col_kernel<HALF, GC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    k2_col_kernel<HALF, GC><<<grid, K2C_THREADS, smem, st>>>(c);
    return cudaGetLastError();
}

int launch_col(wd_handle_t h, const K2Args &a, size_t smem)
{
    K2CArgs c;
    c.p = a.p; c.beta = a.beta; c.nb = a.nb; c.out = a.out;
    c.handoff = a.handoff; c.flags = a.nostore ? K2C_FLAG_NOSTORE : 0;
    c.pair_flag = nullptr;
    // beta <-> pi - beta pairing: decided on the device, no host round trip.  The mirror matrix is computed
    // from the trig values of its partner: |beta[b] + beta[nb-1-b] - pi| <= tol bounds the extra error by j * tol = 2e-13.
    if (a.nb >= 2 * K2C_ANG && h->pairing) {
        int *flag = h->d_pair + (h->pair_idx++ % kPairRing);
        const double tol = 2e-13 / (double)std::max(1, a.p.twoj / 2);
        k2c_pair_check_kernel<<<1, 1024, 0, h->stream>>>(a.beta, a.nb, tol, flag);
        h->launches++;
        c.pair_flag = flag;
    }
    const long long nch = (a.nb + K2C_ANG - 1) / K2C_ANG;
    const long long G = h->num_sms;
    long long items = nch;
    if (nch < G) items = nch * std::max<long long>(1, std::min<long long>(16, G / nch));
    const int grid = (int)std::min<long long>(items, G);
    const int gc = (a.p.Q + 15) >> 4;
    const size_t optin = h->smem_optin;
    cudaStream_t st = h->stream;
    cudaError_t e = cudaErrorInvalidValue;
    if (a.p.twoj & 1) {
        switch (gc) {
        case 1: e = launch_col_gc<true, 1>(c, grid, smem, optin, st); break;
        case 2: e = launch_col_gc<true, 2>(c, grid, smem, optin, st); break;
        case 3: e = launch_col_gc<true, 3>(c, grid, smem, optin, st); break;
        case 4: e = launch_col_gc<true, 4>(c, grid, smem, optin, st); break;
        }
    } else {
        switch (gc) {
        case 1: e = launch_col_gc<false, 1>(c, grid, smem, optin, st); break;
        case 2: e = launch_col_gc<false, 2>(c, grid, smem, optin, st); break;
        case 3: e = launch_col_gc<false, 3>(c, grid, smem, optin, st); break;
        case 4: e = launch_col_gc<false, 4>(c, grid, smem, optin, st); break;
        case 5: e = launch_col_gc<false, 5>(c, grid, smem, optin, st); break;
        case 6: e = launch_col_gc<false, 6>(c, grid, smem, optin, st); break;
        case 7: e = launch_col_gc<false, 7>(c, grid, smem, optin, st); break;
        }
    }
    h->launches++;
    CU(e);
    return WD_OK;
}

// Streaming DMMA kernel (real d, table too large for shared memory): trig table pass + contraction.
template <bool HALF>
int launch_stream(wd_handle_t h, const Plan &cpl, const K2Args &a)
{
    // chunk-major copy of the table, built once per plan (the plan lives in h->plans; the handle's lock is held)
    Plan &pl = const_cast<Plan &>(cpl);
    const int nck0 = k2s_nchunks(a.p.K0p), nck1 = k2s_nchunks(a.p.K1p);
    if (!pl.uqc) {
        const int qpad = k2s_qpad(a.p.Q, HALF);
        const size_t n = (size_t)(nck0 + nck1) * qpad * K2S_LDU;
        double *buf = nullptr;
        if (cudaMalloc(&buf, n * sizeof(double)) != cudaSuccess) return fail(WD_ERR_NOMEM, "cudaMalloc of the chunk-major table of 2j = %d failed", a.p.twoj);
        CU(cudaMemsetAsync(buf, 0, n * sizeof(double), h->stream));
        const long long total = (long long)(nck0 + nck1) * a.p.Q * K2S_KC;
        k2s_table_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 8), 256, 0, h->stream>>>(a.p, buf, qpad, nck0, nck1);
        h->launches++;
        // one-time per plan: finish before the pointer is published (the handle's stream may be changed between calls)
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { cudaFree(buf); return fail(WD_ERR_CUDA, "building the chunk-major table of 2j = %d failed: %s", a.p.twoj, cudaGetErrorString(e)); }
        pl.uqc = buf; pl.uqc_qpad = qpad;
    }
    const K2SSmem L = k2s_smem_layout(HALF);
    const size_t smem = (size_t)L.total * sizeof(double);
    if (!h->attr_set_stream[HALF]) {
        CU(cudaFuncSetAttribute(k2_stream_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
        h->attr_set_stream[HALF] = true;
    }
    K2SArgs sa;
    sa.base = a;
    sa.nck0 = nck0; sa.nck1 = nck1;
    sa.uqc = pl.uqc; sa.qpad = pl.uqc_qpad;
    const long long nac = (a.nb + K2_ANG - 1) / K2_ANG;
    const size_t ntrig = (size_t)nac * (sa.nck0 + sa.nck1) * K2S_TRIG_TILE;
    double *trig = nullptr;
    if (h->pool) CU(cudaMallocFromPoolAsync(&trig, ntrig * sizeof(double), h->pool, h->stream));
    else CU(cudaMallocAsync(&trig, ntrig * sizeof(double), h->stream));
    sa.trig = trig;
    const long long work = nac * (sa.nck0 + sa.nck1) * K2_ANG * K2S_KC;
    const int tblocks = (int)std::min<long long>((work + 255) / 256, (long long)h->num_sms * 16);
    k2s_trig_kernel<<<tblocks, 256, 0, h->stream>>>(a.p, a.beta, a.nb, trig, sa.nck0, sa.nck1);
    const int gmax = HALF ? 2 : 4;
    const int ngroups = (a.p.Q + 15) / 16;
    const long long ntiles = nac * ((ngroups + gmax - 1) / gmax) * ((a.p.Q + K2_CWARPS - 1) / K2_CWARPS);
    const int grid = (int)std::min<long long>(ntiles, h->num_sms);
    k2_stream_kernel<HALF><<<grid, K2_THREADS_REAL, smem, h->stream>>>(sa);
    h->launches += 2;
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(trig, h->stream);
    if (e != cudaSuccess) return fail(WD_ERR_CUDA, "streaming kernel launch failed: %s", cudaGetErrorString(e));
    return WD_OK;
}

// d (or D) for nb generic angles, all pointers on the device.
int contract_device(wd_handle_t h, const Plan &pl, const double *alpha, const double *beta, const double *gamma,
                    long long nb, double *out, bool cplx, bool equator)
{
    if (nb <= 0) return WD_OK;
    K2Args a;
    a.p = pl.dev;
    a.alpha = alpha; a.beta = beta; a.gamma = gamma;
    a.nb = nb; a.out = out; a.equator = equator ? 1 : 0;
    a.nsplit = 1; a.nostore = 0; a.nfull = 0; a.nitems = 0; a.tsplit = 1;
    {
        const char *e = getenv("WD_K2_NOSTORE");         // kernel-experiment knob: 1 = no store instructions (SM-side time)
        if (e) a.nostore = atoi(e) != 0;
    }
    {
        // hand-off between the two compute warps of a sub-partition: the partner is released when ~2/3 of the k-steps of a
        // main loop are done (real d, second-generation kernel, cfg-2: release at 50 % 9.43 ms, 58-73 % 9.21-9.24 ms, 81 %
        // 9.31 ms, 96 % 10.1 ms, off 10.6 ms; complex D, whose warps also drain their tiles: 80 %).
        // Encoding: 0 = off; h = release after h-1 pairs of k-steps of class 0, or (h-1 >= 100) h-101 pairs of class 1.
        const int s0 = pl.dev.K0p / 4, s1 = pl.dev.K1p / 4;
        const int target = cplx ? (4 * (s0 + s1) + 4) / 5 : (2 * (s0 + s1) + 1) / 3;
        a.handoff = 1 + (target <= s0 ? target / 2 : 100 + (target - s0) / 2);
        const char *e = getenv("WD_K2_HANDOFF");        // kernel-experiment knob
        if (e) a.handoff = atoi(e);
    }
    size_t smem = 0;
    const bool fits = cplx ? cplx_fits(h, pl.dev, &smem) : real_fits(h, pl.dev, &smem);
    const int variant = h->variant;
    if (variant == 2 && !fits) return fail(WD_ERR_UNSUPPORTED, "2j = %d does not fit the DMMA kernel (%zu B of shared memory)", pl.dev.twoj, smem);
    // Equator() requests (single matrices, exact zeros of :230-237) take the plain kernel
    if (variant == 3 && !cplx && !equator) return (pl.dev.twoj & 1) ? launch_stream<true>(h, pl, a) : launch_stream<false>(h, pl, a);
    const bool half = pl.dev.twoj & 1;
    if (!cplx && !equator && (variant == 0 || variant == 4)) {
        size_t csm = 0;
        if (col_fits(h, pl.dev, &csm)) return launch_col(h, a, csm);
        if (variant == 4) return fail(WD_ERR_UNSUPPORTED, "2j = %d does not fit the column-staged DMMA kernel (%zu B of shared memory)", pl.dev.twoj, csm);
    }
    const bool use_dmma = !equator && ((variant == 2) || (variant == 0 && fits));
    if (use_dmma && !cplx) return half ? launch_real<true>(h, a, smem) : launch_real<false>(h, a, smem);
    if (use_dmma) return half ? launch_cplx<true>(h, a, smem) : launch_cplx<false>(h, a, smem);
    // real d for a table that does not fit in shared memory: the streaming DMMA kernel
    if (!cplx && !equator && variant != 1) return half ? launch_stream<true>(h, pl, a) : launch_stream<false>(h, pl, a);
    const size_t psm = 2 * (size_t)pl.dev.Kp * sizeof(double);
    const int Q = pl.dev.Q;
    const int gy = std::max(1, std::min(64, (Q * Q + 255) / 256));
    for (long long b0 = 0; b0 < nb; b0 += 1 << 30) {     // grid.x limit
        K2Args s = a;
        const long long cnt = std::min<long long>(nb - b0, 1 << 30);
        s.beta = beta + b0; s.nb = cnt;
        if (cplx) { s.alpha = alpha + b0; s.gamma = gamma + b0; }
        s.out = out + (size_t)b0 * pl.dev.N * pl.dev.N * (cplx ? 2 : 1);
        dim3 grid((unsigned)cnt, gy);
        if (cplx) k2_plain_kernel<true><<<grid, 256, psm, h->stream>>>(s);
        else k2_plain_kernel<false><<<grid, 256, psm, h->stream>>>(s);
        h->launches++;
    }
    CU(cudaGetLastError());
    return WD_OK;
}

// host-memory front end: angles H2D, slabs of output computed on the device and copied back, double-buffered.
int contract_host(wd_handle_t h, const Plan &pl, const double *alpha, const double *beta, const double *gamma,
                  long long nb, double *out, bool cplx, bool equator)
{
    if (nb <= 0) return WD_OK;
    CU(cudaSetDevice(h->device));
    const size_t per = (size_t)pl.dev.N * pl.dev.N * (cplx ? 2 : 1);          // doubles per matrix
    const int nang = cplx ? 3 : 1;
    auto grow = [](double **buf, size_t *have, size_t need) -> cudaError_t {
        if (*have >= need) return cudaSuccess;
        if (*buf) cudaFree(*buf);
        *buf = nullptr; *have = 0;
        cudaError_t e = cudaMalloc(buf, need);
        if (e == cudaSuccess) *have = need;
        return e;
    };
    if (grow(&h->ws_ang, &h->ws_ang_bytes, (size_t)nb * nang * sizeof(double)) != cudaSuccess)
        return fail(WD_ERR_NOMEM, "cudaMalloc of the angle workspace failed");
    double *d_ang = h->ws_ang;
    cudaMemcpyAsync(d_ang, beta, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (cplx) {
        cudaMemcpyAsync(d_ang + nb, alpha, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(d_ang + 2 * nb, gamma, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    }
    const size_t slab_bytes_target = (size_t)512 << 20;
    long long slab = std::max<long long>(1, (long long)(slab_bytes_target / (per * sizeof(double))));
    slab = std::min(slab, nb);
    if (slab > K2_ANG) slab -= slab % K2_ANG;
    const int nbuf = (slab < nb) ? 2 : 1;
    int rc = WD_OK;
    for (int i = 0; i < nbuf && rc == WD_OK; ++i) {
        if (grow(&h->ws_slab[i], &h->ws_slab_bytes[i], (size_t)slab * per * sizeof(double)) != cudaSuccess)
            rc = fail(WD_ERR_NOMEM, "cudaMalloc of a %zu-byte output slab failed", (size_t)slab * per * sizeof(double));
        if (!h->ev_done[i]) {
            cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming);
        }
    }
    int k = 0;
    for (long long b0 = 0; b0 < nb && rc == WD_OK; b0 += slab, ++k) {
        const int i = k % nbuf;
        const long long cnt = std::min(slab, nb - b0);
        if (k >= nbuf) cudaStreamWaitEvent(h->stream, h->ev_copied[i], 0);  // slab buffer free again
        rc = contract_device(h, pl, cplx ? d_ang + nb + b0 : nullptr, d_ang + b0, cplx ? d_ang + 2 * nb + b0 : nullptr,
                             cnt, h->ws_slab[i], cplx, equator);
        if (rc) break;
        cudaEventRecord(h->ev_done[i], h->stream);
        cudaStreamWaitEvent(h->copy_stream, h->ev_done[i], 0);
        cudaMemcpyAsync(out + (size_t)b0 * per, h->ws_slab[i], (size_t)cnt * per * sizeof(double), cudaMemcpyDeviceToHost,
                        h->copy_stream);
        cudaEventRecord(h->ev_copied[i], h->copy_stream);
    }
    cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->copy_stream);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(WD_ERR_CUDA, "contraction failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return WD_OK;
}

__global__ void pole_kernel(double *D, int twoj, int kind, int cplx, double alpha, double gamma)
{
    // wignerd!(…, NorthPole/SouthPole) (src/WignerD.jl:247-264) then, for complex D, the phase loop :338-340
    const int N = twoj + 1;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N * N; e += gridDim.x * blockDim.x) {
        const int row = e % N, col = e / N;
        if (kind == WD_BETA_NORTH && row == col) {
            if (cplx) { D[2 * e] = 1.0; D[2 * e + 1] = 0.0; } else D[e] = 1.0;
        }
        if (kind == WD_BETA_SOUTH && row + col == N - 1) {
            const double v = (row & 1) ? -1.0 : 1.0;       // + at m = -j, alternating
            if (cplx) { D[2 * e] = v; D[2 * e + 1] = 0.0; } else D[e] = v;
        }
        if (cplx) {
            const double m = 0.5 * (double)(2 * row - twoj), n = 0.5 * (double)(2 * col - twoj);
            double s, c;
            sincos(-(m * alpha + n * gamma), &s, &c);
            const double re = D[2 * e], im = D[2 * e + 1];
            D[2 * e] = re * c - im * s;
            D[2 * e + 1] = re * s + im * c;
        }
    }
}

// marks present[2j] = 1 for every 2j of the tuple list; present[kMaxTwoJ + 1] collects bit 1 (a negative j) and bit 2 (j too large)
__global__ void mark_twoj_kernel(const int32_t *v, long long n, int *present, int max_twoj)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = v[i];
        if (t < 0) atomicOr(present + max_twoj + 1, 1);
        else if (t > max_twoj) atomicOr(present + max_twoj + 1, 2);
        else if (present[t] == 0) present[t] = 1;
    }
}

int ensure_table(wd_handle_t h, int max_twoj)
{
    if (!h->table_dirty && h->table_max >= max_twoj) return WD_OK;
    int top = -1;
    for (auto &kv : h->plans) top = std::max(top, kv.first);
    std::vector<PlanDev> tab((size_t)top + 1);
    for (auto &t : tab) { memset(&t, 0, sizeof t); }
    for (auto &kv : h->plans) tab[kv.first] = kv.second.dev;
    if (h->d_table) cudaFree(h->d_table);
    h->d_table = nullptr;
    CU(cudaMalloc(&h->d_table, std::max<size_t>(1, tab.size()) * sizeof(PlanDev)));
    CU(cudaMemcpy(h->d_table, tab.data(), tab.size() * sizeof(PlanDev), cudaMemcpyHostToDevice));
    h->table_max = top;
    h->table_dirty = false;
    return WD_OK;
}

int jmn_common(wd_handle_t h, const int32_t *twoj, const int32_t *twom, const int32_t *twon, const double *alpha,
               const double *beta, const double *gamma, int64_t n, int beta_kind, double *out, int mem, bool cplx)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (n < 0) return fail(WD_ERR_ASSERT, "negative tuple count");
    if (n == 0) return WD_OK;
    if (!twoj || !twom || !twon || !out || (beta_kind == WD_BETA_GENERIC && !beta) || (cplx && (!alpha || !gamma)))
        return fail(WD_ERR_ASSERT, "null pointer argument");
    if (beta_kind < 0 || beta_kind > 3) return fail(WD_ERR_ASSERT, "bad beta_kind %d", beta_kind);
    std::lock_guard<std::mutex> lk(h->mtx);
    CU(cudaSetDevice(h->device));
    const int32_t *d_tj = twoj, *d_tm = twom, *d_tn = twon;
    const double *d_a = alpha, *d_b = beta, *d_g = gamma;
    double *d_out = out;
    void *blk = nullptr;
    const size_t osz = (cplx ? 2 : 1) * sizeof(double) * (size_t)n;
    if (mem == WD_MEM_HOST) {
        const size_t isz = sizeof(int32_t) * (size_t)n, dsz = sizeof(double) * (size_t)n;
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };      // every sub-buffer 256-byte aligned
        const size_t total = 3 * al(isz) + 3 * al(dsz) + al(osz);
        CU(cudaMalloc(&blk, total));
        char *pch = (char *)blk;
        double *pb = (double *)pch; pch += al(dsz);
        double *pa = (double *)pch; pch += al(dsz);
        double *pg = (double *)pch; pch += al(dsz);
        double *po = (double *)pch; pch += al(osz);
        int32_t *pj = (int32_t *)pch; pch += al(isz);
        int32_t *pm = (int32_t *)pch; pch += al(isz);
        int32_t *pn = (int32_t *)pch;
        cudaMemcpyAsync(pj, twoj, isz, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(pm, twom, isz, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(pn, twon, isz, cudaMemcpyHostToDevice, h->stream);
        if (beta) cudaMemcpyAsync(pb, beta, dsz, cudaMemcpyHostToDevice, h->stream);
        if (cplx) {
            cudaMemcpyAsync(pa, alpha, dsz, cudaMemcpyHostToDevice, h->stream);
            cudaMemcpyAsync(pg, gamma, dsz, cudaMemcpyHostToDevice, h->stream);
        }
        d_tj = pj; d_tm = pm; d_tn = pn; d_a = pa; d_b = pb; d_g = pg; d_out = po;
    }
    int rc = WD_OK;
    int hflag[2] = {0, 0};
    do {
        // poles never touch the eigenvectors (src/WignerD.jl:218-228): no tables needed
        int maxtj = -1;
        if (beta_kind == WD_BETA_GENERIC || beta_kind == WD_BETA_EQUATOR) {
            // only the j values that occur are decomposed and cached, like the reference's Jy_eigen (:163-176)
            cudaMemsetAsync(h->d_flag, 0, 2 * sizeof(int), h->stream);
            cudaMemsetAsync(h->d_present, 0, (kMaxTwoJ + 2) * sizeof(int), h->stream);
            mark_twoj_kernel<<<(int)std::min<long long>(1024, (n + 255) / 256), 256, 0, h->stream>>>(d_tj, n, h->d_present, kMaxTwoJ);
            h->launches++;
            std::vector<int> present((size_t)kMaxTwoJ + 2);
            if (cudaMemcpyAsync(present.data(), h->d_present, present.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = fail(WD_ERR_CUDA, "scan of the j values failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
            if (present[kMaxTwoJ + 1] & 1) { rc = fail(WD_ERR_ASSERT, "j must be >= 0"); break; }
            if (present[kMaxTwoJ + 1] & 2) { rc = fail(WD_ERR_UNSUPPORTED, "2j exceeds the supported maximum %d", kMaxTwoJ); break; }
            std::vector<int32_t> all;
            for (int i = 0; i <= kMaxTwoJ; ++i) if (present[i]) { all.push_back(i); maxtj = i; }
            rc = prepare_locked(h, all.data(), (int)all.size());
            if (rc) break;
            rc = ensure_table(h, maxtj);
            if (rc) break;
        } else {
            cudaMemsetAsync(h->d_flag, 0, 2 * sizeof(int), h->stream);
        }
        K4Args a;
        a.plans = h->d_table; a.max_twoj = maxtj;
        a.twoj = d_tj; a.twom = d_tm; a.twon = d_tn;
        a.beta = d_b; a.alpha = d_a; a.gamma = d_g;
        a.n = n; a.beta_kind = beta_kind; a.out = d_out; a.err = h->d_flag;
        const int blocks = (int)std::min<long long>((n + 7) / 8, (long long)h->num_sms * 8);
        if (cplx) k4_jmn_kernel<true><<<blocks, 256, 0, h->stream>>>(a);
        else k4_jmn_kernel<false><<<blocks, 256, 0, h->stream>>>(a);
        h->launches++;
        if (mem == WD_MEM_HOST) cudaMemcpyAsync(out, d_out, osz, cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(hflag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { rc = fail(WD_ERR_CUDA, "jmn kernel failed: %s", cudaGetErrorString(e)); break; }
        if (hflag[0]) rc = fail(WD_ERR_ARGUMENT, "m or n is inconsistent with j for at least one tuple");
    } while (0);
    if (blk) cudaFree(blk);
    return rc;
}

}  // namespace

// ======================================================================================================
// C ABI
// ======================================================================================================
extern "C" {

const char *wd_last_error(void) { return g_err.c_str(); }
int wd_abi_version(void) { return 1; }

int wd_create(int device, wd_handle_t *out)
{
    if (!out) return fail(WD_ERR_ASSERT, "null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(WD_ERR_CUDA, "no CUDA device is usable (%s); this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(WD_ERR_ASSERT, "device %d out of range (0..%d)", device, ndev - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(WD_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    wd_handle_t h = new wd_handle_s;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CU(cudaMalloc(&h->d_flag, 4 * sizeof(int)));
    CU(cudaMalloc(&h->d_pair, kPairRing * sizeof(int)));
    CU(cudaMalloc(&h->d_present, (kMaxTwoJ + 2) * sizeof(int)));
    CU(cudaMemset(h->d_pair, 0, kPairRing * sizeof(int)));
    {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&h->pool, &props) == cudaSuccess) {
            uint64_t never = UINT64_MAX;
            cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &never);
        } else {
            h->pool = nullptr;
            cudaGetLastError();                           // not fatal: fall back to the default pool
        }
    }
    *out = h;
    return WD_OK;
}

int wd_destroy(wd_handle_t h)
{
    if (!h) return WD_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto &kv : h->plans) { cudaFree(kv.second.uq); if (kv.second.uqc) cudaFree(kv.second.uqc); }
    if (h->d_table) cudaFree(h->d_table);
    if (h->d_flag) cudaFree(h->d_flag);
    if (h->d_pair) cudaFree(h->d_pair);
    if (h->d_present) cudaFree(h->d_present);
    if (h->ws_ang) cudaFree(h->ws_ang);
    for (int i = 0; i < 2; ++i) {
        if (h->ws_slab[i]) cudaFree(h->ws_slab[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    }
    if (h->pool) cudaMemPoolDestroy(h->pool);
    cudaStreamDestroy(h->own_stream);
    cudaStreamDestroy(h->copy_stream);
    delete h;
    return WD_OK;
}

int wd_set_stream(wd_handle_t h, void *s)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    std::lock_guard<std::mutex> lk(h->mtx);
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return WD_OK;
}

int wd_synchronize(wd_handle_t h)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->copy_stream));
    return WD_OK;
}

int64_t wd_launch_count(wd_handle_t h) { return h ? h->launches : 0; }

int wd_cache_clear(wd_handle_t h)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    std::lock_guard<std::mutex> lk(h->mtx);
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto &kv : h->plans) { cudaFree(kv.second.uq); if (kv.second.uqc) cudaFree(kv.second.uqc); }
    h->plans.clear();
    h->table_dirty = true;
    if (h->pool) cudaMemPoolTrimTo(h->pool, 0);
    return WD_OK;
}

int wd_set_variant(wd_handle_t h, int variant)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (variant < 0 || variant > 5) return fail(WD_ERR_ASSERT, "variant must be 0..5");
    std::lock_guard<std::mutex> lk(h->mtx);
    h->pairing = (variant == 5) ? 0 : 1;                 // 5 = auto without the beta <-> pi - beta pairing
    h->variant = (variant == 5) ? 0 : variant;
    return WD_OK;
}

int wd_prepare(wd_handle_t h, const int32_t *twoj_list, int32_t nj)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (nj < 0 || (nj > 0 && !twoj_list)) return fail(WD_ERR_ASSERT, "bad j list");
    std::lock_guard<std::mutex> lk(h->mtx);
    return prepare_locked(h, twoj_list, nj);
}

int wd_eigvecs(wd_handle_t h, int32_t twoj, double *u_out)
{
    if (!h || !u_out) return fail(WD_ERR_ASSERT, "null argument");
    std::lock_guard<std::mutex> lk(h->mtx);
    const Plan *pl;
    int rc = get_plan(h, twoj, &pl);
    if (rc) return rc;
    const PlanDev &d = pl->dev;
    std::vector<double> uq((size_t)d.Q * d.ldu);
    CU(cudaMemcpy(uq.data(), pl->uq, uq.size() * sizeof(double), cudaMemcpyDeviceToHost));
    const int N = d.N, K = d.Q, i0 = N - d.Q, k0 = N - K;
    for (int kappa = 0; kappa < K; ++kappa) {
        const int kp = wd_kp_of_kappa(kappa, K, d.K0p);
        const double eps = (kp < d.K0p) ? 1.0 : -1.0;        // u[-m] = eps u[m]
        const int k = k0 + kappa, kmir = N - 1 - k;          // mu and -mu columns
        for (int i = 0; i < N; ++i) {
            const double v = (i >= i0) ? uq[(size_t)(i - i0) * d.ldu + kp] : eps * uq[(size_t)(N - 1 - i - i0) * d.ldu + kp];
            u_out[i + (size_t)N * k] = v;
            if (kmir != k) u_out[i + (size_t)N * kmir] = (i & 1) ? -v : v;    // u[i,-mu] = (-1)^i u[i,mu]
        }
    }
    return WD_OK;
}

static int matrices(wd_handle_t h, int32_t twoj, const double *alpha, const double *beta, const double *gamma,
                    int64_t nb, double *out, int mem, bool cplx, bool equator)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (nb < 0) return fail(WD_ERR_ASSERT, "negative batch size");
    if (nb == 0) return WD_OK;
    if (!beta || !out || (cplx && (!alpha || !gamma))) return fail(WD_ERR_ASSERT, "null pointer argument");
    if (mem != WD_MEM_HOST && mem != WD_MEM_DEVICE) return fail(WD_ERR_ASSERT, "bad mem flag %d", mem);
    std::lock_guard<std::mutex> lk(h->mtx);
    const Plan *pl;
    int rc = get_plan(h, twoj, &pl);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    if (mem == WD_MEM_DEVICE) return contract_device(h, *pl, alpha, beta, gamma, nb, out, cplx, equator);
    return contract_host(h, *pl, alpha, beta, gamma, nb, out, cplx, equator);
}

int wd_d_batch(wd_handle_t h, int32_t twoj, const double *beta, int64_t nb, double *out, int mem)
{
    return matrices(h, twoj, nullptr, beta, nullptr, nb, out, mem, false, false);
}

int wd_D_batch(wd_handle_t h, int32_t twoj, const double *alpha, const double *beta, const double *gamma, int64_t nb,
               double *out, int mem)
{
    return matrices(h, twoj, alpha, beta, gamma, nb, out, mem, true, false);
}

static int pole_matrix(wd_handle_t h, int32_t twoj, int kind, bool cplx, double alpha, double gamma, double *D)
{
    // the eigen step still runs first in the reference (:306-307); keep the cache behaviour identical
    std::lock_guard<std::mutex> lk(h->mtx);
    const Plan *pl;
    int rc = get_plan(h, twoj, &pl);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    const size_t bytes = (size_t)pl->dev.N * pl->dev.N * (cplx ? 2 : 1) * sizeof(double);
    double *d_D = nullptr;
    CU(cudaMalloc(&d_D, bytes));
    cudaMemcpyAsync(d_D, D, bytes, cudaMemcpyHostToDevice, h->stream);
    pole_kernel<<<std::max(1, std::min(64, (pl->dev.N * pl->dev.N + 255) / 256)), 256, 0, h->stream>>>(d_D, twoj, kind, cplx ? 1 : 0, alpha, gamma);
    h->launches++;
    cudaMemcpyAsync(D, d_D, bytes, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(d_D);
    if (e != cudaSuccess) return fail(WD_ERR_CUDA, "pole fill failed: %s", cudaGetErrorString(e));
    return WD_OK;
}

int wd_d_matrix(wd_handle_t h, int32_t twoj, double beta, int beta_kind, double *d)
{
    if (!h || !d) return fail(WD_ERR_ASSERT, "null argument");
    if (twoj < 0) return fail(WD_ERR_ASSERT, "j must be >= 0");
    switch (beta_kind) {
    case WD_BETA_GENERIC: return matrices(h, twoj, nullptr, &beta, nullptr, 1, d, WD_MEM_HOST, false, false);
    case WD_BETA_EQUATOR: { const double b = 1.5707963267948966; return matrices(h, twoj, nullptr, &b, nullptr, 1, d, WD_MEM_HOST, false, true); }
    case WD_BETA_NORTH:
    case WD_BETA_SOUTH: return pole_matrix(h, twoj, beta_kind, false, 0.0, 0.0, d);
    default: return fail(WD_ERR_ASSERT, "bad beta_kind %d", beta_kind);
    }
}

int wd_D_matrix(wd_handle_t h, int32_t twoj, double alpha, double beta, int beta_kind, double gamma, double *D)
{
    if (!h || !D) return fail(WD_ERR_ASSERT, "null argument");
    if (twoj < 0) return fail(WD_ERR_ASSERT, "j must be >= 0");
    switch (beta_kind) {
    case WD_BETA_GENERIC: return matrices(h, twoj, &alpha, &beta, &gamma, 1, D, WD_MEM_HOST, true, false);
    case WD_BETA_EQUATOR: { const double b = 1.5707963267948966; return matrices(h, twoj, &alpha, &b, &gamma, 1, D, WD_MEM_HOST, true, true); }
    case WD_BETA_NORTH:
    case WD_BETA_SOUTH: return pole_matrix(h, twoj, beta_kind, true, alpha, gamma, D);
    default: return fail(WD_ERR_ASSERT, "bad beta_kind %d", beta_kind);
    }
}

int wd_d_multi(wd_handle_t h, const int32_t *twoj_list, int32_t nj, const double *beta, int64_t nb, double *out,
               const int64_t *offsets, int mem)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (nj < 0 || nb < 0) return fail(WD_ERR_ASSERT, "negative count");
    if (nj == 0 || nb == 0) return WD_OK;
    if (!twoj_list || !beta || !out || !offsets) return fail(WD_ERR_ASSERT, "null pointer argument");
    std::lock_guard<std::mutex> lk(h->mtx);
    int rc = prepare_locked(h, twoj_list, nj);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    if (mem == WD_MEM_DEVICE) {
        for (int k = 0; k < nj; ++k) {
            rc = contract_device(h, h->plans[twoj_list[k]], nullptr, beta, nullptr, nb, out + offsets[k], false, false);
            if (rc) return rc;
        }
        return WD_OK;
    }
    for (int k = 0; k < nj; ++k) {
        rc = contract_host(h, h->plans[twoj_list[k]], nullptr, beta, nullptr, nb, out + offsets[k], false, false);
        if (rc) return rc;
    }
    return WD_OK;
}

int wd_djmn_batch(wd_handle_t h, const int32_t *twoj, const int32_t *twom, const int32_t *twon, const double *beta,
                  int64_t n, int beta_kind, double *out, int mem)
{
    return jmn_common(h, twoj, twom, twon, nullptr, beta, nullptr, n, beta_kind, out, mem, false);
}

int wd_Djmn_batch(wd_handle_t h, const int32_t *twoj, const int32_t *twom, const int32_t *twon, const double *alpha,
                  const double *beta, const double *gamma, int64_t n, int beta_kind, double *out, int mem)
{
    return jmn_common(h, twoj, twom, twon, alpha, beta, gamma, n, beta_kind, out, mem, true);
}

int64_t wd_k2_shared_bytes(int32_t twoj, int kernel)
{
    if (twoj < 0 || twoj > kMaxTwoJ) return -1;
    PlanDev p{};
    plan_dims(twoj, p);
    switch (kernel) {
    case 0: return (int64_t)k2r_smem_layout(p.Q, p.Kp, p.ldu).total * (int64_t)sizeof(double);
    case 1: return (int64_t)k2_smem_layout(p.Q, p.Kp, p.ldu).total * (int64_t)sizeof(double);
    case 2: return (int64_t)k2s_smem_layout((twoj & 1) != 0).total * (int64_t)sizeof(double);
    default: return -1;
    }
}

int wd_fp64_peak(wd_handle_t h, int kind, int iters, double *tflops)
{
    if (!h || !tflops) return fail(WD_ERR_ASSERT, "null argument");
    std::lock_guard<std::mutex> lk(h->mtx);
    CU(cudaSetDevice(h->device));
    double *sink = nullptr;
    CU(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = h->num_sms * 4, threads = 256;
    fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(kind, 16, sink);   // warm-up
    cudaEventRecord(e0, h->stream);
    fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>(kind, iters, sink);
    cudaEventRecord(e1, h->stream);
    h->launches += 2;
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess) return fail(WD_ERR_CUDA, "peak kernel failed: %s", cudaGetErrorString(e));
    const double warps = (double)blocks * threads / 32.0;
    const double flops = (kind == 0) ? warps * iters * 8.0 * 512.0 : (double)blocks * threads * iters * 16.0 * 2.0;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return WD_OK;
}

}  // extern "C"
