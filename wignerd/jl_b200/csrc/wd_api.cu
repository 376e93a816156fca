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
#include "wd_k3_tma.cuh"
#include "wd_rotate.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <emmintrin.h>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
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

// [lo, hi) of the contiguous slab of n batch items owned by rank (of world): slab starts are multiples of `align` items (the
// kernels' angle-chunk size), sizes differ by at most `align`, the slabs partition [0, n) exactly.  Mirrors shard.py.
void shard_bounds(long long n, int rank, int world, int align, long long *lo, long long *hi)
{
    const long long nchunks = (n + align - 1) / align;
    const long long base = nchunks / world, rem = nchunks % world;
    const long long lo_c = rank * base + std::min<long long>(rank, rem);
    const long long hi_c = lo_c + base + (rank < rem ? 1 : 0);
    *lo = std::min(lo_c * align, n);
    *hi = std::min(hi_c * align, n);
}

// host copy with non-temporal stores: the destination (a pageable output array, written once) is not read for ownership and
// does not evict the bounce buffer from the caches
void stream_copy(char *dst, const char *src, size_t n)
{
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) != 0) { memcpy(dst, src, n); return; }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_load_si128((const __m128i *)(src + i)), b = _mm_load_si128((const __m128i *)(src + i + 16));
        const __m128i c = _mm_load_si128((const __m128i *)(src + i + 32)), d = _mm_load_si128((const __m128i *)(src + i + 48));
        _mm_stream_si128((__m128i *)(dst + i), a); _mm_stream_si128((__m128i *)(dst + i + 16), b);
        _mm_stream_si128((__m128i *)(dst + i + 32), c); _mm_stream_si128((__m128i *)(dst + i + 48), d);
    }
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
}

// A few persistent host threads that copy one block of memory in parallel (pinned bounce buffer -> pageable destination).
class CopyPool {
public:
    explicit CopyPool(int nthreads) : n_(std::max(1, nthreads))
    {
        for (int i = 0; i < n_; ++i) th_.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void copy(void *dst, const void *src, size_t bytes)
    {
        if (bytes < ((size_t)1 << 20)) { memcpy(dst, src, bytes); return; }
        std::unique_lock<std::mutex> lk(m_);
        dst_ = (char *)dst; src_ = (const char *)src; bytes_ = bytes; left_ = n_; ++gen_;
        cv_.notify_all();
        done_.wait(lk, [this] { return left_ == 0; });
    }
private:
    void loop(int i)
    {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (stop_) return;
            char *d = dst_; const char *s = src_; const size_t b = bytes_;
            lk.unlock();
            const size_t per = ((b / n_) + 4095) & ~(size_t)4095;
            const size_t lo = std::min(b, per * i), hi = std::min(b, per * (i + 1));
            if (hi > lo) stream_copy(d + lo, s + lo, hi - lo);
            lk.lock();
            if (--left_ == 0) done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    unsigned long long gen_ = 0;
    bool stop_ = false;
    char *dst_ = nullptr; const char *src_ = nullptr; size_t bytes_ = 0; int left_ = 0;
};

struct Plan {
    PlanDev dev{};
    double *uq = nullptr, *mu = nullptr, *w = nullptr;
    double *uqc = nullptr;          // chunk-major copy of uq for the streaming kernel, built on first use
    int uqc_qpad = 0;
    double *ufull = nullptr;        // full N x N eigenvector matrix (column-major, ld = N) for wd_rotate_sh, built on first use
};

}  // namespace

struct wd_handle_s {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    std::mutex mtx;
    std::map<int, Plan> plans;
    std::vector<void *> table_blocks;   // device allocations that hold the plans' tables (one per eigensolver launch / loaded record)
    int num_sms = 0;
    size_t smem_optin = 0;
    int64_t launches = 0;
    int variant = 0;
    int pairing = 1;                 // beta <-> pi - beta pairing in k2_col_kernel (wd_set_variant 6 switches it off)
    size_t slab_bytes = (size_t)512 << 20;   // host path: output slab per kernel launch / device-to-host copy (wd_set_slab_bytes)
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
    bool attr_set_col[2][8] = {};
    bool attr_set_k3t[2][8] = {};
    int *d_pair = nullptr;            // ring of device flags written by k2c_pair_check_kernel (one per call in flight)
    unsigned pair_idx = 0;
    const int *shared_pair_flag = nullptr;   // wd_d_multi: ONE symmetry check of the angle grid (at the tightest tolerance) for all j
    // host-path workspace, kept across calls (grow-only): angle vectors and two output slabs + their events
    double *ws_ang = nullptr, *ws_slab[2] = {nullptr, nullptr};
    size_t ws_ang_bytes = 0, ws_slab_bytes[2] = {0, 0};
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    // pageable host destinations: ring of pinned bounce buffers + host copy threads (created on first use)
    static constexpr int kBounce = 4;
    static constexpr size_t kBounceBytes = (size_t)32 << 20;
    char *bounce[kBounce] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_bounce[kBounce] = {nullptr, nullptr, nullptr, nullptr};
    CopyPool *pool_copy = nullptr;
    // scalar calls (one matrix, a handful of tuples): a small device block and its pinned mirror, allocated once
    static constexpr size_t kSmallBytes = (size_t)4 << 20;
    char *small_dev = nullptr, *small_host = nullptr;
    // multi-GPU handle (wd_create_multi): one single-device child handle per entry; the parent owns no device state
    std::vector<wd_handle_t> children;
};

namespace {

int for_each_child(wd_handle_t h, const std::function<int(wd_handle_t, int)> &f);

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
    // ONE device block for the tables of the whole batch (a cudaMalloc per plan made "every 2j <= 1024" cost between 0.16 and
    // 2.5 s depending on the box), 256-byte aligned sub-blocks: the kernels bring UQ in with bulk-TMA loads
    size_t tables_total = 0;
    std::vector<size_t> toff(n);
    for (size_t t = 0; t < n; ++t) {
        const PlanDev &d = fresh[t].dev;
        toff[t] = tables_total;
        tables_total += ((size_t)d.Q * d.ldu + 2 * (size_t)d.Kp + 31) & ~(size_t)31;
    }
    double *scratch = nullptr, *block = nullptr;
    EigJob *d_jobs = nullptr;
    auto cleanup = [&](bool drop_tables) {
        if (scratch) cudaFree(scratch);
        if (d_jobs) cudaFree(d_jobs);
        if (drop_tables && block) cudaFree(block);
    };
    cudaError_t e = cudaMalloc(&scratch, scratch_total * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_jobs, n * sizeof(EigJob));
    if (e == cudaSuccess) e = cudaMalloc(&block, tables_total * sizeof(double));
    if (e != cudaSuccess) { cleanup(true); return fail(WD_ERR_NOMEM, "cudaMalloc of the eigensolver workspace and tables (%zu bytes) failed: %s", (scratch_total + tables_total) * sizeof(double), cudaGetErrorString(e)); }
    cudaMemsetAsync(block, 0, tables_total * sizeof(double), h->stream);
    size_t off = 0;
    for (size_t t = 0; t < n; ++t) {
        Plan &pl = fresh[t];
        const PlanDev &d = pl.dev;
        const size_t nuq = (size_t)d.Q * d.ldu;
        pl.uq = block + toff[t];
        pl.mu = pl.uq + nuq;
        pl.w = pl.mu + d.Kp;
        pl.dev.uq = pl.uq; pl.dev.mu = pl.mu; pl.dev.w = pl.w;
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
    h->table_blocks.push_back(block);
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
#ifdef WD_EXPERIMENT
    {
        const char *e = getenv("WD_K2_TAIL");             // kernel-experiment knob: 0 = never split the last wave
        if (e && atoi(e) == 0 && b.nfull > 0) b.tsplit = 1;
    }
#endif
    b.nitems = b.nfull + rem * b.tsplit;
    const int grid = (int)std::min<long long>(b.nitems, G);
    k2_real_kernel<HALF><<<grid, K2_THREADS_REAL, smem, h->stream>>>(b);
    h->launches++;
    CU(cudaGetLastError());
    return WD_OK;
}

// Column-staged DMMA kernel for real d (k2_col_kernel, wd_k2_col.cuh): the default for tables that fit in shared memory
constexpr int kPairRing = 64;
// beta <-> pi - beta pairing: the mirror matrix is computed from the trig values of its partner, so a grid counts as symmetric
// when |beta[b] + beta[nb-1-b] - pi| <= tol for every b; |d d/d beta| <= j bounds the extra error of the mirror matrix by
// j * tol <= 5e-13 (2e-13 up to j = 100; a linspace grid deviates by <= 8e-16).
double pair_tol(int twoj) { return std::min(2e-15, 5e-13 / (double)std::max(1, twoj / 2)); }

// the device flag that says whether beta[0..nb) is symmetric about pi/2: the one wd_d_multi made for the whole call, or a fresh check
const int *pair_flag_for(wd_handle_t h, const double *beta, long long nb, int twoj)
{
    if (h->shared_pair_flag) return h->shared_pair_flag;
    int *flag = h->d_pair + (h->pair_idx++ % kPairRing);
    k2c_pair_check_kernel<<<1, 1024, 0, h->stream>>>(beta, nb, pair_tol(twoj), flag);
    h->launches++;
    return flag;
}
bool col_fits(wd_handle_t h, const PlanDev &p, size_t *bytes)
{
    const K2CSmem L = k2c_smem_layout(p.Q, p.Kp, p.ldu, p.N);
    *bytes = (size_t)L.total * sizeof(double);
    return k2c_shape_ok(p.twoj) && *bytes <= h->smem_optin;
}

template <bool HALF, int GC>
cudaError_t launch_col_gc(wd_handle_t h, const K2CArgs &c, int grid, size_t smem)
{
    bool &attr_set = h->attr_set_col[HALF][GC];           // per handle = per device: the attribute belongs to the device's context
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k2_col_kernel<HALF, GC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    k2_col_kernel<HALF, GC><<<grid, K2C_THREADS, smem, h->stream>>>(c);
    return cudaGetLastError();
}

// mode 0: the column kernel works only if the device-side pair check finds the grid symmetric (the caller launches
// k2_real_kernel with skip_flag = *flag_out alongside); 1: column kernel for every grid, paired when symmetric;
// 2: column kernel, never paired
int launch_col(wd_handle_t h, const K2Args &a, size_t smem, int mode, const int **flag_out)
{
    K2CArgs c;
    c.p = a.p; c.beta = a.beta; c.nb = a.nb; c.out = a.out;
    c.flags = a.nostore ? K2C_FLAG_NOSTORE : 0;
    c.pair_flag = nullptr;
    c.require_pair = mode == 0;
    {
        // hand-off at ~80 % of the k-steps (cfg-2: 50 % 6.73 ms, 65 % 6.50, 81 % 6.39, 100 % 6.64, off 6.90 ms)
        const int s0 = a.p.K0p / 4, s1 = a.p.K1p / 4;
        const int target = (4 * (s0 + s1) + 4) / 5;
        c.handoff = 1 + (target <= s0 ? target / 2 : 100 + (target - s0) / 2);
#ifdef WD_EXPERIMENT
        const char *e = getenv("WD_K2_HANDOFF");
        if (e) c.handoff = atoi(e);
#endif
    }
    // beta <-> pi - beta pairing: decided on the device, no host round trip (pair_tol)
    if (mode != 2) {
        c.pair_flag = pair_flag_for(h, a.beta, a.nb, a.p.twoj);
        if (flag_out) *flag_out = c.pair_flag;
    }
    const long long nch = (a.nb + K2C_ANG - 1) / K2C_ANG;
    const long long G = h->num_sms;
    long long items = nch;
    if (nch < G) items = nch * std::max<long long>(1, std::min<long long>(16, G / nch));
    const int grid = (int)std::min<long long>(items, G);
    const int gc = (a.p.Q + 15) >> 4;
    cudaError_t e = cudaErrorInvalidValue;
    if (a.p.twoj & 1) {
        switch (gc) {
        case 1: e = launch_col_gc<true, 1>(h, c, grid, smem); break;
        case 2: e = launch_col_gc<true, 2>(h, c, grid, smem); break;
        case 3: e = launch_col_gc<true, 3>(h, c, grid, smem); break;
        case 4: e = launch_col_gc<true, 4>(h, c, grid, smem); break;
        }
    } else {
        switch (gc) {
        case 1: e = launch_col_gc<false, 1>(h, c, grid, smem); break;
        case 2: e = launch_col_gc<false, 2>(h, c, grid, smem); break;
        case 3: e = launch_col_gc<false, 3>(h, c, grid, smem); break;
        case 4: e = launch_col_gc<false, 4>(h, c, grid, smem); break;
        case 5: e = launch_col_gc<false, 5>(h, c, grid, smem); break;
        case 6: e = launch_col_gc<false, 6>(h, c, grid, smem); break;
        case 7: e = launch_col_gc<false, 7>(h, c, grid, smem); break;
        }
    }
    h->launches++;
    CU(e);
    return WD_OK;
}

// Tensor-store DMMA kernel for complex D (k3_tma_kernel, wd_k3_tma.cuh): the default for complex D when its boxes fit
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime (the library links the static runtime only, not libcuda)
EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return (EncodeTiledFn)f;
    }();
    return fn;
}

constexpr int kK3TWarps = 8;         // two warps per SM sub-partition (three were tried)
bool k3t_fits(wd_handle_t h, const PlanDev &p, const double *out, size_t *bytes)
{
    *bytes = (size_t)k3t_smem_layout(p.Q, p.Kp, p.ldu, kK3TWarps).total * sizeof(double);
    return k3t_shape_ok(p.twoj) && *bytes <= h->smem_optin && (((uintptr_t)out) & 15) == 0 && encode_tiled_fn() != nullptr;
}

template <bool HALF, int GC>
cudaError_t launch_k3t_gc(wd_handle_t h, const K3TArgs &c, const CUtensorMap &ta, const CUtensorMap &td, int grid, size_t smem)
{
    bool &attr_set = h->attr_set_k3t[HALF][GC];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k3_tma_kernel<HALF, GC, kK3TWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    k3_tma_kernel<HALF, GC, kK3TWarps><<<grid, kK3TWarps * 32, smem, h->stream>>>(c, ta, td);
    return cudaGetLastError();
}
int launch_k3t(wd_handle_t h, const K2Args &a, size_t smem)
{
    K3TArgs c;
    c.p = a.p; c.alpha = a.alpha; c.beta = a.beta; c.gamma = a.gamma; c.nb = a.nb;
    c.flags = a.nostore ? K3T_FLAG_NOSTORE : 0;
    c.handoff = a.handoff;
    // the output as a 2-d tensor of doubles: dim0 = the 2 N^2 doubles of one matrix, dim1 = the batch; boxes of 8 angles x half a column
    const int N = a.p.N, Q = a.p.Q, i0 = N - Q;
    CUtensorMap tm[2];
    const cuuint64_t dims[2] = {(cuuint64_t)2 * N * N, (cuuint64_t)a.nb};
    const cuuint64_t strides[1] = {(cuuint64_t)N * N * 16};
    const cuuint32_t estr[2] = {1, 1};
    for (int t = 0; t < 2; ++t) {
        const cuuint32_t box[2] = {(cuuint32_t)(2 * (t == 0 ? Q : i0)), (cuuint32_t)K3T_ANG};
        const CUresult r = encode_tiled_fn()(&tm[t], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)a.out, dims, strides, box, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(WD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for 2j = %d, %lld matrices", (int)r, a.p.twoj, a.nb);
    }
    const long long nch = (a.nb + K3T_ANG - 1) / K3T_ANG;
    const long long G = h->num_sms;
    long long items = nch;
    if (nch < G) items = nch * std::max<long long>(1, std::min<long long>(16, G / nch));
    const int grid = (int)std::min<long long>(items, G);
    const int gc = (Q + 15) >> 4;
    cudaError_t e = cudaErrorInvalidValue;
    if (a.p.twoj & 1) {
        switch (gc) {
        case 1: e = launch_k3t_gc<true, 1>(h, c, tm[0], tm[1], grid, smem); break;
        case 2: e = launch_k3t_gc<true, 2>(h, c, tm[0], tm[1], grid, smem); break;
        case 3: e = launch_k3t_gc<true, 3>(h, c, tm[0], tm[1], grid, smem); break;
        case 4: e = launch_k3t_gc<true, 4>(h, c, tm[0], tm[1], grid, smem); break;
        }
    } else {
        switch (gc) {
        case 1: e = launch_k3t_gc<false, 1>(h, c, tm[0], tm[1], grid, smem); break;
        case 2: e = launch_k3t_gc<false, 2>(h, c, tm[0], tm[1], grid, smem); break;
        case 3: e = launch_k3t_gc<false, 3>(h, c, tm[0], tm[1], grid, smem); break;
        case 4: e = launch_k3t_gc<false, 4>(h, c, tm[0], tm[1], grid, smem); break;
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
    // the per-call trig table ([angle chunk][mu chunk] tiles of 9 KB) is bounded: batches are cut into angle slabs of at most
    // 256 MB of table (1.0 GB for j = 512 on 10^5 angles, 10 GB on 10^6, otherwise)
    const int nck = nck0 + nck1;
    const long long nac_all = (a.nb + K2_ANG - 1) / K2_ANG;
    const long long nac_cap = std::max<long long>(1, (((long long)256 << 20) / (long long)sizeof(double)) / ((long long)nck * K2S_TRIG_TILE));
    for (long long ac0 = 0; ac0 < nac_all; ac0 += nac_cap) {
        const long long nac = std::min(nac_cap, nac_all - ac0);
        const long long b0 = ac0 * K2_ANG, cnt = std::min<long long>(nac * K2_ANG, a.nb - b0);
        K2SArgs sa;
        sa.base = a;
        sa.base.beta = a.beta + b0; sa.base.nb = cnt;
        sa.base.out = a.out + (size_t)b0 * a.p.N * a.p.N;
        sa.nck0 = nck0; sa.nck1 = nck1;
        sa.uqc = pl.uqc; sa.qpad = pl.uqc_qpad;
        // symmetric grids: pairing, decided on the device (only when the batch is not cut into slabs)
        sa.pair_flag = nullptr;
        if (h->pairing && nac_all <= nac_cap && a.nb >= 32) sa.pair_flag = pair_flag_for(h, a.beta, a.nb, a.p.twoj);
        const size_t ntrig = (size_t)nac * nck * K2S_TRIG_TILE;
        double *trig = nullptr;
        if (h->pool) CU(cudaMallocFromPoolAsync(&trig, ntrig * sizeof(double), h->pool, h->stream));
        else CU(cudaMallocAsync(&trig, ntrig * sizeof(double), h->stream));
        sa.trig = trig;
        const long long work = nac * nck * K2_ANG * K2S_KC;
        const int tblocks = (int)std::min<long long>((work + 255) / 256, (long long)h->num_sms * 16);
        k2s_trig_kernel<<<tblocks, 256, 0, h->stream>>>(a.p, sa.base.beta, cnt, trig, sa.nck0, sa.nck1, sa.pair_flag);
        const int gmax = HALF ? 2 : 4;
        const int ngroups = (a.p.Q + 15) / 16;
        const long long ntiles = nac * ((ngroups + gmax - 1) / gmax) * ((a.p.Q + K2_CWARPS - 1) / K2_CWARPS);
        const int grid = (int)std::min<long long>(ntiles, h->num_sms);
        k2_stream_kernel<HALF><<<grid, K2_THREADS_REAL, smem, h->stream>>>(sa);
        h->launches += 2;
        cudaError_t e = cudaGetLastError();
        cudaFreeAsync(trig, h->stream);
        if (e != cudaSuccess) return fail(WD_ERR_CUDA, "streaming kernel launch failed: %s", cudaGetErrorString(e));
    }
    return WD_OK;
}

// D[b][n][m] = d[b][n][m] * cis(-(m alpha_b + n gamma_b)): the phase loop of wignerD! (src/WignerD.jl:338-340), one element per thread
__global__ void __launch_bounds__(256) phase_kernel(const double *__restrict__ d, const double *__restrict__ alpha, const double *__restrict__ gamma,
                                                    long long nb, int twoj, double2 *__restrict__ D)
{
    const int N = twoj + 1;
    const long long NN = (long long)N * N, total = nb * NN;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / NN;
        const int r = (int)(e - b * NN), row = r % N, col = r / N;
        const double m = 0.5 * (double)(2 * row - twoj), n = 0.5 * (double)(2 * col - twoj);
        double s, c;
        sincos(-(m * alpha[b] + n * gamma[b]), &s, &c);
        const double v = d[e];
        D[e] = make_double2(v * c, v * s);
    }
}

// d (or D) for nb generic angles, all pointers on the device.
int contract_device(wd_handle_t h, const Plan &pl, const double *alpha, const double *beta, const double *gamma,
                    long long nb, double *out, bool cplx, bool equator)
{
    if (nb <= 0) return WD_OK;
    // complex matrices are written with 16-byte stores (and as a TMA tensor): complex128 arrays are always aligned like that
    if (cplx && (((uintptr_t)out) & 15) != 0) return fail(WD_ERR_ASSERT, "complex output must be 16-byte aligned (got %p)", (void *)out);
    K2Args a;
    a.p = pl.dev;
    a.alpha = alpha; a.beta = beta; a.gamma = gamma;
    a.nb = nb; a.out = out; a.equator = equator ? 1 : 0;
    a.nsplit = 1; a.nostore = 0; a.nfull = 0; a.nitems = 0; a.tsplit = 1; a.skip_flag = nullptr;
#ifdef WD_EXPERIMENT
    {
        const char *e = getenv("WD_K2_NOSTORE");         // kernel-experiment knob: 1 = no store instructions (SM-side time)
        if (e) a.nostore = atoi(e) != 0;
    }
#endif
    {
        // hand-off between the two compute warps of a sub-partition: the partner is released when ~2/3 of the k-steps of a
        // main loop are done (real d, second-generation kernel, cfg-2: release at 50 % 9.43 ms, 58-73 % 9.21-9.24 ms, 81 %
        // 9.31 ms, 96 % 10.1 ms, off 10.6 ms; complex D, whose warps also drain their tiles: 80 %).
        // Encoding: 0 = off; h = release after h-1 pairs of k-steps of class 0, or (h-1 >= 100) h-101 pairs of class 1.
        const int s0 = pl.dev.K0p / 4, s1 = pl.dev.K1p / 4;
        const int target = cplx ? (4 * (s0 + s1) + 4) / 5 : (2 * (s0 + s1) + 1) / 3;
        a.handoff = 1 + (target <= s0 ? target / 2 : 100 + (target - s0) / 2);
#ifdef WD_EXPERIMENT
        const char *e = getenv("WD_K2_HANDOFF");        // kernel-experiment knob
        if (e) a.handoff = atoi(e);
#endif
    }
    size_t smem = 0;
    const bool fits = cplx ? cplx_fits(h, pl.dev, &smem) : real_fits(h, pl.dev, &smem);
    // variants 4 / 5 select the real-d column kernel only: complex D and Equator requests are dispatched as in auto mode
    // variant 7 selects the complex-D tensor-store kernel only: real d is dispatched as in auto mode
    const int variant = (((cplx || equator) && (h->variant == 4 || h->variant == 5)) || ((!cplx || equator) && h->variant == 7)) ? 0 : h->variant;
    if (variant == 2 && !fits) return fail(WD_ERR_UNSUPPORTED, "2j = %d does not fit the DMMA kernel (%zu B of shared memory)", pl.dev.twoj, smem);
    // Equator() requests (single matrices, exact zeros of :230-237) take the plain kernel
    if (variant == 3 && !cplx && !equator) return (pl.dev.twoj & 1) ? launch_stream<true>(h, pl, a) : launch_stream<false>(h, pl, a);
    const bool half = pl.dev.twoj & 1;
    if (!cplx && !equator && (variant == 0 || variant == 4 || variant == 5)) {
        size_t csm = 0;
        const bool cfits = col_fits(h, pl.dev, &csm);
        if (variant != 0) {
            if (!cfits) return fail(WD_ERR_UNSUPPORTED, "2j = %d does not fit the column-staged DMMA kernel (%zu B of shared memory)", pl.dev.twoj, csm);
            return launch_col(h, a, csm, variant == 4 ? 1 : 2, nullptr);
        }
        // auto: symmetric grids (beta[b] + beta[nb-1-b] = pi) -> column kernel with pairing, everything else -> k2_real_kernel.
        // Both are launched; the device-side check lets exactly one of them work (an empty launch costs ~3 us).
        // (small batches are launch-latency bound -- 17 us at cfg-1 -- and stay with the one kernel)
        if (cfits && fits && h->pairing && (double)nb * pl.dev.N * pl.dev.N >= 16.0e6) {
            const int *flag = nullptr;
            int rc = launch_col(h, a, csm, 0, &flag);
            if (rc) return rc;
            a.skip_flag = flag;
            return half ? launch_real<true>(h, a, smem) : launch_real<false>(h, a, smem);
        }
    }
    if (cplx && !equator && (variant == 0 || variant == 7)) {
        // complex D: the tensor-store kernel (k3_tma_kernel) when its boxes fit in shared memory, else k2_cplx_kernel / streaming
        size_t tsm = 0;
        const bool tfits = k3t_fits(h, pl.dev, out, &tsm);
        if (tfits) return launch_k3t(h, a, tsm);
        if (variant == 7) return fail(WD_ERR_UNSUPPORTED, "2j = %d (or the alignment of the output) does not fit the tensor-store kernel (%zu B of shared memory)", pl.dev.twoj, tsm);
    }
    const bool use_dmma = !equator && ((variant == 2) || ((variant == 0 || variant == 7) && fits));
    if (use_dmma && !cplx) return half ? launch_real<true>(h, a, smem) : launch_real<false>(h, a, smem);
    if (use_dmma) return half ? launch_cplx<true>(h, a, smem) : launch_cplx<false>(h, a, smem);
    // real d for a table that does not fit in shared memory: the streaming DMMA kernel
    if (!cplx && !equator && variant != 1) return half ? launch_stream<true>(h, pl, a) : launch_stream<false>(h, pl, a);
    if (cplx && !equator && variant != 1) {
        // complex D beyond the resident kernel (2j > 223): real d of a slab of angles through the streaming DMMA kernel into a
        // temporary, then one pass that applies cis(-(m alpha + n gamma)) exactly as src/WignerD.jl:339 does (the reference's own
        // two-step form).  The phase pass moves 24 bytes per element; at j = 512 that is a third of the contraction's time,
        // against the FP64-ALU kernel's O(N^3) multiply-adds (~20 x slower than the tensor path).
        const size_t per = (size_t)pl.dev.N * pl.dev.N;
        const long long slab = std::max<long long>(K2_ANG, std::min<long long>(nb, (((long long)1 << 30) / (long long)(per * sizeof(double))) / K2_ANG * K2_ANG));
        double *tmp = nullptr;
        if (h->pool) CU(cudaMallocFromPoolAsync(&tmp, (size_t)slab * per * sizeof(double), h->pool, h->stream));
        else CU(cudaMallocAsync(&tmp, (size_t)slab * per * sizeof(double), h->stream));
        int rc = WD_OK;
        for (long long b0 = 0; b0 < nb && rc == WD_OK; b0 += slab) {
            const long long cnt = std::min(slab, nb - b0);
            K2Args s = a;
            s.beta = beta + b0; s.nb = cnt; s.out = tmp; s.alpha = nullptr; s.gamma = nullptr;
            rc = half ? launch_stream<true>(h, pl, s) : launch_stream<false>(h, pl, s);
            if (rc) break;
            const long long total = cnt * (long long)per;
            phase_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 32), 256, 0, h->stream>>>(
                tmp, alpha + b0, gamma + b0, cnt, pl.dev.twoj, reinterpret_cast<double2 *>(out) + (size_t)b0 * per);
            h->launches++;
        }
        cudaFreeAsync(tmp, h->stream);
        if (rc) return rc;
        CU(cudaGetLastError());
        return WD_OK;
    }
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

// pinned (or registered / managed) host memory?  Pageable memory makes cudaMemcpyAsync a staged, synchronous copy.
bool host_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

int ensure_bounce(wd_handle_t h)
{
    if (h->pool_copy) return WD_OK;
    for (int i = 0; i < wd_handle_s::kBounce; ++i) {
        if (cudaMallocHost((void **)&h->bounce[i], wd_handle_s::kBounceBytes) != cudaSuccess)
            return fail(WD_ERR_NOMEM, "cudaMallocHost of a %zu-byte bounce buffer failed", wd_handle_s::kBounceBytes);
        cudaEventCreateWithFlags(&h->ev_bounce[i], cudaEventDisableTiming);
    }
    const unsigned hw = std::thread::hardware_concurrency();
    h->pool_copy = new CopyPool((int)std::max(2u, std::min(12u, hw > 4 ? hw - 4 : hw / 2)));
    return WD_OK;
}

// device -> host copy of one finished slab on the copy stream.  Pinned destination: one cudaMemcpyAsync.  Pageable
// destination: 32 MB pieces through a ring of pinned bounce buffers, each piece moved on by the host copy threads while the
// next pieces are in flight (a plain cudaMemcpyAsync into pageable memory is staged by the driver through one small buffer
// and serialises with the host).
int copy_out(wd_handle_t h, char *dst, const char *src_dev, size_t bytes, bool pinned)
{
    if (pinned) {
        cudaError_t e = cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, h->copy_stream);
        if (e != cudaSuccess) return fail(WD_ERR_CUDA, "device-to-host copy failed: %s", cudaGetErrorString(e));
        return WD_OK;
    }
    int rc = ensure_bounce(h);
    if (rc) return rc;
    constexpr int R = wd_handle_s::kBounce;
    const size_t P = wd_handle_s::kBounceBytes;
    const size_t npieces = (bytes + P - 1) / P;
    size_t issued = 0, drained = 0;
    while (drained < npieces) {
        while (issued < npieces && issued - drained < (size_t)R) {
            const size_t off = issued * P, len = std::min(P, bytes - off);
            cudaMemcpyAsync(h->bounce[issued % R], src_dev + off, len, cudaMemcpyDeviceToHost, h->copy_stream);
            cudaEventRecord(h->ev_bounce[issued % R], h->copy_stream);
            ++issued;
        }
        cudaError_t e = cudaEventSynchronize(h->ev_bounce[drained % R]);
        if (e != cudaSuccess) return fail(WD_ERR_CUDA, "device-to-host copy failed: %s", cudaGetErrorString(e));
        const size_t off = drained * P, len = std::min(P, bytes - off);
        h->pool_copy->copy(dst + off, h->bounce[drained % R], len);
        ++drained;
    }
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
    const size_t total_bytes = (size_t)nb * per * sizeof(double);
    // small pageable outputs go through the driver's own staging; large ones through the bounce ring
    const bool pinned = total_bytes < ((size_t)8 << 20) || host_is_pinned(out);
    const size_t slab_bytes_target = h->slab_bytes;
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
        rc = copy_out(h, (char *)(out + (size_t)b0 * per), (const char *)h->ws_slab[i], (size_t)cnt * per * sizeof(double), pinned);
        cudaEventRecord(h->ev_copied[i], h->copy_stream);
    }
    cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->copy_stream);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(WD_ERR_CUDA, "contraction failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return WD_OK;
}

// Rotation of spherical-harmonic coefficients, all pointers on the device (wd_rotate.cuh): per l one phase kernel and two
// tensor-core GEMMs with the full eigenvector matrix of l, on the handle's stream.
int rotate_device(wd_handle_t h, int lmax, const double *alm, long long alm_stride, const double *alpha, const double *beta,
                  const double *gamma, long long nb, double *out)
{
    if (nb <= 0) return WD_OK;
    const long long L2 = (long long)(lmax + 1) * (lmax + 1);
    const int Nmax = 2 * lmax + 1;
    double *ws = nullptr;                                  // c and z: Nmax x nb complex each
    const size_t wsn = (size_t)Nmax * nb * 2;
    if (h->pool) CU(cudaMallocFromPoolAsync(&ws, 2 * wsn * sizeof(double), h->pool, h->stream));
    else CU(cudaMallocAsync(&ws, 2 * wsn * sizeof(double), h->stream));
    double *cbuf = ws, *zbuf = ws + wsn;
    int rc = WD_OK;
    for (int l = lmax; l >= 0 && rc == WD_OK; --l) {
        Plan &pl = h->plans[2 * l];
        const int N = 2 * l + 1;
        if (!pl.ufull) {
            double *u = nullptr;
            if (cudaMalloc(&u, (size_t)N * N * sizeof(double)) != cudaSuccess) { rc = fail(WD_ERR_NOMEM, "cudaMalloc of the full eigenvector matrix of l = %d failed", l); break; }
            k5_expand_u_kernel<<<(int)std::min<long long>(((long long)N * N + 255) / 256, (long long)h->num_sms * 8), 256, 0, h->stream>>>(pl.dev, u, N);
            h->launches++;
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);       // one-time per l: finish before the pointer is published
            if (e != cudaSuccess) { cudaFree(u); rc = fail(WD_ERR_CUDA, "expanding the eigenvectors of l = %d failed: %s", l, cudaGetErrorString(e)); break; }
            pl.ufull = u;
        }
        const long long work = (long long)N * nb;
        k5_prephase_kernel<<<(int)std::min<long long>((work + 255) / 256, (long long)h->num_sms * 16), 256, 0, h->stream>>>(
            reinterpret_cast<const double2 *>(alm), alm_stride, l, alpha, nb, reinterpret_cast<double2 *>(cbuf));
        K5Args g;
        g.U = pl.ufull; g.ld = N; g.N = N; g.l = l; g.nb = nb; g.out_stride = L2;
        // 64-column tiles when 128-column tiles would leave SMs without a CTA (two CTAs fit per SM)
        const unsigned rt = (unsigned)((N + K5_TM - 1) / K5_TM);
        const bool narrow = (long long)rt * ((2 * nb + 127) / 128) < 2LL * h->num_sms;
        const int TN = narrow ? 64 : 128;
        dim3 grid((unsigned)((2 * nb + TN - 1) / TN), rt);
        g.in = cbuf; g.out = zbuf; g.angle = beta;
        if (narrow) k5_gemm_kernel<true, 64><<<grid, 128, 0, h->stream>>>(g); else k5_gemm_kernel<true, 128><<<grid, 256, 0, h->stream>>>(g);
        g.in = zbuf; g.out = out; g.angle = gamma;
        if (narrow) k5_gemm_kernel<false, 64><<<grid, 128, 0, h->stream>>>(g); else k5_gemm_kernel<false, 128><<<grid, 256, 0, h->stream>>>(g);
        h->launches += 3;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(WD_ERR_CUDA, "rotation kernels of l = %d failed to launch: %s", l, cudaGetErrorString(e));
    }
    cudaFreeAsync(ws, h->stream);
    return rc;
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
    if (cplx && mem == WD_MEM_DEVICE && (((uintptr_t)out) & 15) != 0) return fail(WD_ERR_ASSERT, "complex output must be 16-byte aligned (got %p)", (void *)out);
    if (!h->children.empty()) {
        if (mem != WD_MEM_HOST) return fail(WD_ERR_UNSUPPORTED, "a multi-GPU handle takes host buffers; for device buffers use the per-device handles (wd_multi_child)");
        const int R = (int)h->children.size();
        return for_each_child(h, [=](wd_handle_t c, int r) {
            long long lo, hi;
            shard_bounds(n, r, R, 1, &lo, &hi);
            if (hi <= lo) return (int)WD_OK;
            return jmn_common(c, twoj + lo, twom + lo, twon + lo, alpha ? alpha + lo : nullptr, beta ? beta + lo : nullptr,
                              gamma ? gamma + lo : nullptr, hi - lo, beta_kind, out + (size_t)lo * (cplx ? 2 : 1), mem, cplx);
        });
    }
    std::lock_guard<std::mutex> lk(h->mtx);
    CU(cudaSetDevice(h->device));
    const int32_t *d_tj = twoj, *d_tm = twom, *d_tn = twon;
    const double *d_a = alpha, *d_b = beta, *d_g = gamma;
    double *d_out = out;
    void *blk = nullptr;
    char *stage = nullptr;                      // pinned mirror of the device block (small calls): one copy in, one copy out
    size_t stage_in = 0, stage_out_off = 0;
    const size_t osz = (cplx ? 2 : 1) * sizeof(double) * (size_t)n;
    if (mem == WD_MEM_HOST) {
        const size_t isz = sizeof(int32_t) * (size_t)n, dsz = sizeof(double) * (size_t)n;
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };      // every sub-buffer 256-byte aligned
        const size_t total = 3 * al(isz) + 3 * al(dsz) + al(osz);
        char *base;
        if (total <= wd_handle_s::kSmallBytes) {
            base = h->small_dev; stage = h->small_host;                    // scalar calls: per-handle workspace
        } else {
            CU(cudaMalloc(&blk, total));
            base = (char *)blk;
        }
        char *pch = base;
        double *pb = (double *)pch; pch += al(dsz);
        double *pa = (double *)pch; pch += al(dsz);
        double *pg = (double *)pch; pch += al(dsz);
        int32_t *pj = (int32_t *)pch; pch += al(isz);
        int32_t *pm = (int32_t *)pch; pch += al(isz);
        int32_t *pn = (int32_t *)pch; pch += al(isz);
        double *po = (double *)pch;
        if (stage) {
            memcpy(stage + ((char *)pj - base), twoj, isz);
            memcpy(stage + ((char *)pm - base), twom, isz);
            memcpy(stage + ((char *)pn - base), twon, isz);
            if (beta) memcpy(stage + ((char *)pb - base), beta, dsz);
            if (cplx) { memcpy(stage + ((char *)pa - base), alpha, dsz); memcpy(stage + ((char *)pg - base), gamma, dsz); }
            stage_in = (size_t)((char *)po - base);
            stage_out_off = stage_in;
            cudaMemcpyAsync(base, stage, stage_in, cudaMemcpyHostToDevice, h->stream);
        } else {
            cudaMemcpyAsync(pj, twoj, isz, cudaMemcpyHostToDevice, h->stream);
            cudaMemcpyAsync(pm, twom, isz, cudaMemcpyHostToDevice, h->stream);
            cudaMemcpyAsync(pn, twon, isz, cudaMemcpyHostToDevice, h->stream);
            if (beta) cudaMemcpyAsync(pb, beta, dsz, cudaMemcpyHostToDevice, h->stream);
            if (cplx) {
                cudaMemcpyAsync(pa, alpha, dsz, cudaMemcpyHostToDevice, h->stream);
                cudaMemcpyAsync(pg, gamma, dsz, cudaMemcpyHostToDevice, h->stream);
            }
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
            std::vector<int> present((size_t)kMaxTwoJ + 2, 0);
            if (mem == WD_MEM_HOST) {
                // host tuples: scan them here (a scalar call then needs one synchronisation, not two)
                for (int64_t i = 0; i < n; ++i) {
                    const int t = twoj[i];
                    if (t < 0) present[kMaxTwoJ + 1] |= 1; else if (t > kMaxTwoJ) present[kMaxTwoJ + 1] |= 2; else present[t] = 1;
                }
            } else {
                cudaMemsetAsync(h->d_present, 0, (kMaxTwoJ + 2) * sizeof(int), h->stream);
                mark_twoj_kernel<<<(int)std::min<long long>(1024, (n + 255) / 256), 256, 0, h->stream>>>(d_tj, n, h->d_present, kMaxTwoJ);
                h->launches++;
                if (cudaMemcpyAsync(present.data(), h->d_present, present.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                    cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = fail(WD_ERR_CUDA, "scan of the j values failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
            }
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
        const int per_block = K4_THREADS / K4_G;
        const int blocks = (int)std::min<long long>((n + per_block - 1) / per_block, (long long)h->num_sms * 16);
        if (cplx) k4_jmn_kernel<true><<<blocks, K4_THREADS, 0, h->stream>>>(a);
        else k4_jmn_kernel<false><<<blocks, K4_THREADS, 0, h->stream>>>(a);
        h->launches++;
        if (mem == WD_MEM_HOST) cudaMemcpyAsync(stage ? (void *)(stage + stage_out_off) : (void *)out, d_out, osz, cudaMemcpyDeviceToHost, h->stream);
        cudaMemcpyAsync(hflag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { rc = fail(WD_ERR_CUDA, "jmn kernel failed: %s", cudaGetErrorString(e)); break; }
        if (stage) memcpy(out, stage + stage_out_off, osz);
        if (hflag[0]) rc = fail(WD_ERR_ARGUMENT, "m or n is inconsistent with j for at least one tuple");
    } while (0);
    if (blk) cudaFree(blk);
    return rc;
}



// multi-GPU handle: f(child, rank) runs on one host thread per child; the first error (and its message) is returned
int for_each_child(wd_handle_t h, const std::function<int(wd_handle_t, int)> &f)
{
    const int R = (int)h->children.size();
    std::vector<int> rcs(R, WD_OK);
    std::vector<std::string> msgs(R);
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r)
        th.emplace_back([&, r] {
            rcs[r] = f(h->children[r], r);
            if (rcs[r]) msgs[r] = g_err;                   // thread-local message of the worker
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < R; ++r)
        if (rcs[r]) { g_err = "device " + std::to_string(h->children[r]->device) + ": " + msgs[r]; return rcs[r]; }
    return WD_OK;
}

}  // namespace

// ======================================================================================================
// C ABI
// ======================================================================================================
extern "C" {

const char *wd_last_error(void) { return g_err.c_str(); }
int wd_abi_version(void) { return 1; }

static void destroy_single(wd_handle_t h)
{
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto &kv : h->plans) { if (kv.second.uqc) cudaFree(kv.second.uqc); if (kv.second.ufull) cudaFree(kv.second.ufull); }
    for (void *b : h->table_blocks) cudaFree(b);
    h->table_blocks.clear();
    if (h->d_table) cudaFree(h->d_table);
    if (h->d_flag) cudaFree(h->d_flag);
    if (h->d_pair) cudaFree(h->d_pair);
    if (h->d_present) cudaFree(h->d_present);
    if (h->ws_ang) cudaFree(h->ws_ang);
    if (h->small_dev) cudaFree(h->small_dev);
    if (h->small_host) cudaFreeHost(h->small_host);
    for (int i = 0; i < 2; ++i) {
        if (h->ws_slab[i]) cudaFree(h->ws_slab[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    }
    for (int i = 0; i < wd_handle_s::kBounce; ++i) {
        if (h->bounce[i]) cudaFreeHost(h->bounce[i]);
        if (h->ev_bounce[i]) cudaEventDestroy(h->ev_bounce[i]);
    }
    delete h->pool_copy;
    if (h->pool) cudaMemPoolDestroy(h->pool);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
}

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
    // a failure below releases what has been built so far
    e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    h->stream = h->own_stream;
    if (e == cudaSuccess) e = cudaMalloc(&h->d_flag, 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_pair, kPairRing * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_present, (kMaxTwoJ + 2) * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->d_pair, 0, kPairRing * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->small_dev, wd_handle_s::kSmallBytes);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->small_host, wd_handle_s::kSmallBytes);
    if (e != cudaSuccess) {
        destroy_single(h);
        return fail(e == cudaErrorMemoryAllocation ? WD_ERR_NOMEM : WD_ERR_CUDA, "wd_create: %s", cudaGetErrorString(e));
    }
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

int wd_create_multi(int ndev, const int *devices, wd_handle_t *out)
{
    if (!out) return fail(WD_ERR_ASSERT, "null output pointer");
    *out = nullptr;
    if (ndev < 1 || ndev > 64) return fail(WD_ERR_ASSERT, "ndev must be 1..64 (got %d)", ndev);
    wd_handle_t h = new wd_handle_s;
    for (int i = 0; i < ndev; ++i) {
        wd_handle_t c = nullptr;
        const int rc = wd_create(devices ? devices[i] : i, &c);          // devices == NULL: 0 .. ndev-1; a device may repeat
        if (rc) {
            for (wd_handle_t d : h->children) destroy_single(d);
            delete h;
            return rc;
        }
        h->children.push_back(c);
    }
    h->device = h->children[0]->device;
    *out = h;
    return WD_OK;
}

int wd_multi_size(wd_handle_t h) { return h ? (h->children.empty() ? 1 : (int)h->children.size()) : 0; }

int wd_multi_child(wd_handle_t h, int i, wd_handle_t *child)
{
    if (!h || !child) return fail(WD_ERR_ASSERT, "null argument");
    if (h->children.empty()) { if (i != 0) return fail(WD_ERR_ASSERT, "child %d of a single-device handle", i); *child = h; return WD_OK; }
    if (i < 0 || i >= (int)h->children.size()) return fail(WD_ERR_ASSERT, "child %d out of range (0..%zu)", i, h->children.size() - 1);
    *child = h->children[i];
    return WD_OK;
}

int wd_shard_bounds(int64_t n, int rank, int world, int align, int64_t *lo, int64_t *hi)
{
    if (!lo || !hi) return fail(WD_ERR_ASSERT, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(WD_ERR_ASSERT, "bad rank/world %d/%d", rank, world);
    if (n < 0 || align < 1) return fail(WD_ERR_ASSERT, "negative batch size or bad alignment");
    long long a, b;
    shard_bounds(n, rank, world, align, &a, &b);
    *lo = a; *hi = b;
    return WD_OK;
}

int wd_destroy(wd_handle_t h)
{
    if (!h) return WD_OK;
    if (!h->children.empty()) {
        for (wd_handle_t c : h->children) destroy_single(c);
        delete h;
        return WD_OK;
    }
    destroy_single(h);
    return WD_OK;
}

int wd_set_stream(wd_handle_t h, void *s)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (!h->children.empty()) return fail(WD_ERR_UNSUPPORTED, "a multi-GPU handle has one stream per device: use wd_multi_child");
    std::lock_guard<std::mutex> lk(h->mtx);
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return WD_OK;
}

int wd_synchronize(wd_handle_t h)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (!h->children.empty()) return for_each_child(h, [](wd_handle_t c, int) { return wd_synchronize(c); });
    std::lock_guard<std::mutex> lk(h->mtx);
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->copy_stream));
    return WD_OK;
}

int64_t wd_launch_count(wd_handle_t h)
{
    if (!h) return 0;
    if (!h->children.empty()) { int64_t n = 0; for (wd_handle_t c : h->children) n += wd_launch_count(c); return n; }
    std::lock_guard<std::mutex> lk(h->mtx);
    return h->launches;
}

int wd_cache_clear(wd_handle_t h)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (!h->children.empty()) return for_each_child(h, [](wd_handle_t c, int) { return wd_cache_clear(c); });
    std::lock_guard<std::mutex> lk(h->mtx);
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto &kv : h->plans) { if (kv.second.uqc) cudaFree(kv.second.uqc); if (kv.second.ufull) cudaFree(kv.second.ufull); }
    for (void *b : h->table_blocks) cudaFree(b);
    h->table_blocks.clear();
    h->plans.clear();
    h->table_dirty = true;
    if (h->pool) cudaMemPoolTrimTo(h->pool, 0);
    return WD_OK;
}

int wd_set_variant(wd_handle_t h, int variant)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (variant < 0 || variant > 7) return fail(WD_ERR_ASSERT, "variant must be 0..7");
    if (!h->children.empty()) return for_each_child(h, [variant](wd_handle_t c, int) { return wd_set_variant(c, variant); });
    std::lock_guard<std::mutex> lk(h->mtx);
    h->pairing = variant != 6;                            // 6 = auto without the beta <-> pi - beta pairing of symmetric grids
    h->variant = variant == 6 ? 0 : variant;
    return WD_OK;
}

int wd_prepare(wd_handle_t h, const int32_t *twoj_list, int32_t nj)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (nj < 0 || (nj > 0 && !twoj_list)) return fail(WD_ERR_ASSERT, "bad j list");
    if (!h->children.empty()) return for_each_child(h, [=](wd_handle_t c, int) { return wd_prepare(c, twoj_list, nj); });
    std::lock_guard<std::mutex> lk(h->mtx);
    return prepare_locked(h, twoj_list, nj);
}

int wd_eigvecs(wd_handle_t h, int32_t twoj, double *u_out)
{
    if (!h || !u_out) return fail(WD_ERR_ASSERT, "null argument");
    if (!h->children.empty()) h = h->children[0];
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
    if (!h->children.empty()) {
        // multi-GPU handle: contiguous 16-aligned slabs of the angle batch per device, no inter-GPU traffic
        if (mem != WD_MEM_HOST) return fail(WD_ERR_UNSUPPORTED, "a multi-GPU handle takes host buffers; for device buffers use the per-device handles (wd_multi_child) with wd_shard_bounds");
        if (twoj < 0) return fail(WD_ERR_ASSERT, "j must be >= 0 (got 2j = %d)", twoj);
        const int R = (int)h->children.size();
        const size_t per = (size_t)(twoj + 1) * (twoj + 1) * (cplx ? 2 : 1);
        return for_each_child(h, [=](wd_handle_t c, int r) {
            long long lo, hi;
            shard_bounds(nb, r, R, K2_ANG, &lo, &hi);
            if (hi <= lo) return (int)WD_OK;
            return matrices(c, twoj, cplx ? alpha + lo : nullptr, beta + lo, cplx ? gamma + lo : nullptr, hi - lo, out + (size_t)lo * per, mem, cplx, equator);
        });
    }
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
    if (!h->children.empty()) h = h->children[0];
    std::lock_guard<std::mutex> lk(h->mtx);
    const Plan *pl;
    int rc = get_plan(h, twoj, &pl);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    const size_t bytes = (size_t)pl->dev.N * pl->dev.N * (cplx ? 2 : 1) * sizeof(double);
    const bool small = bytes <= wd_handle_s::kSmallBytes;
    double *d_D = small ? (double *)h->small_dev : nullptr;
    if (!small) CU(cudaMalloc(&d_D, bytes));
    if (small) { memcpy(h->small_host, D, bytes); cudaMemcpyAsync(d_D, h->small_host, bytes, cudaMemcpyHostToDevice, h->stream); }
    else cudaMemcpyAsync(d_D, D, bytes, cudaMemcpyHostToDevice, h->stream);
    pole_kernel<<<std::max(1, std::min(64, (pl->dev.N * pl->dev.N + 255) / 256)), 256, 0, h->stream>>>(d_D, twoj, kind, cplx ? 1 : 0, alpha, gamma);
    h->launches++;
    cudaMemcpyAsync(small ? (void *)h->small_host : (void *)D, d_D, bytes, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (small) memcpy(D, h->small_host, bytes); else cudaFree(d_D);
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
    if (!h->children.empty()) {
        // every device computes all j for its slab of the angle grid (cfg-4: 4096 / G angles per GPU): matrices [lo, hi) of j_k
        // live at out + offsets[k] + lo * N_k^2
        if (mem != WD_MEM_HOST) return fail(WD_ERR_UNSUPPORTED, "a multi-GPU handle takes host buffers; for device buffers use the per-device handles (wd_multi_child) with wd_shard_bounds");
        for (int k = 0; k < nj; ++k) if (twoj_list[k] < 0) return fail(WD_ERR_ASSERT, "j must be >= 0 (got 2j = %d)", twoj_list[k]);
        const int R = (int)h->children.size();
        return for_each_child(h, [=](wd_handle_t c, int r) {
            long long lo, hi;
            shard_bounds(nb, r, R, K2_ANG, &lo, &hi);
            if (hi <= lo) return (int)WD_OK;
            std::vector<int64_t> off(nj);
            for (int k = 0; k < nj; ++k) off[k] = offsets[k] + (int64_t)lo * (twoj_list[k] + 1) * (twoj_list[k] + 1);
            return wd_d_multi(c, twoj_list, nj, beta + lo, hi - lo, out, off.data(), mem);
        });
    }
    std::lock_guard<std::mutex> lk(h->mtx);
    int rc = prepare_locked(h, twoj_list, nj);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    if (mem == WD_MEM_DEVICE) {
        // one symmetry check of the grid for the whole list, at the tolerance of the largest j
        int tjmax = 0;
        for (int k = 0; k < nj; ++k) tjmax = std::max(tjmax, twoj_list[k]);
        if (h->pairing && nb >= 32) h->shared_pair_flag = pair_flag_for(h, beta, nb, tjmax);
        for (int k = 0; k < nj && rc == WD_OK; ++k)
            rc = contract_device(h, h->plans[twoj_list[k]], nullptr, beta, nullptr, nb, out + offsets[k], false, false);
        h->shared_pair_flag = nullptr;
        return rc;
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

int wd_rotate_sh(wd_handle_t h, int32_t lmax, const double *alm, int64_t alm_stride, const double *alpha, const double *beta,
                 const double *gamma, int64_t nb, double *out, int mem)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (lmax < 0) return fail(WD_ERR_ASSERT, "lmax must be >= 0");
    if (2 * (int64_t)lmax > kMaxTwoJ) return fail(WD_ERR_UNSUPPORTED, "lmax = %d exceeds the supported maximum %d", lmax, kMaxTwoJ / 2);
    if (nb < 0) return fail(WD_ERR_ASSERT, "negative batch size");
    if (nb == 0) return WD_OK;
    if (!alm || !alpha || !beta || !gamma || !out) return fail(WD_ERR_ASSERT, "null pointer argument");
    if (mem != WD_MEM_HOST && mem != WD_MEM_DEVICE) return fail(WD_ERR_ASSERT, "bad mem flag %d", mem);
    const int64_t L2 = (int64_t)(lmax + 1) * (lmax + 1);
    if (alm_stride != 0 && alm_stride < L2) return fail(WD_ERR_ASSERT, "alm_stride must be 0 (one shared set) or >= (lmax+1)^2 = %lld", (long long)L2);
    if (!h->children.empty()) {
        if (mem != WD_MEM_HOST) return fail(WD_ERR_UNSUPPORTED, "a multi-GPU handle takes host buffers; for device buffers use the per-device handles (wd_multi_child)");
        const int R = (int)h->children.size();
        return for_each_child(h, [=](wd_handle_t c, int r) {
            long long lo, hi;
            shard_bounds(nb, r, R, 1, &lo, &hi);
            if (hi <= lo) return (int)WD_OK;
            return wd_rotate_sh(c, lmax, alm + 2 * (size_t)lo * alm_stride, alm_stride, alpha + lo, beta + lo, gamma + lo, hi - lo,
                                out + 2 * (size_t)lo * L2, mem);
        });
    }
    std::lock_guard<std::mutex> lk(h->mtx);
    std::vector<int32_t> js((size_t)lmax + 1);
    for (int l = 0; l <= lmax; ++l) js[l] = 2 * l;
    int rc = prepare_locked(h, js.data(), (int)js.size());
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    if (mem == WD_MEM_DEVICE) return rotate_device(h, lmax, alm, alm_stride, alpha, beta, gamma, nb, out);
    // host buffers: coefficients and angles in, rotated coefficients out, in slabs of rotations (two output slabs in flight)
    const size_t per = 2 * (size_t)L2;                                            // doubles per rotation
    const size_t alm_doubles = alm_stride ? 2 * (size_t)alm_stride * nb : per;
    double *d_in = nullptr;
    CU(cudaMalloc(&d_in, (alm_doubles + 3 * (size_t)nb) * sizeof(double)));
    double *d_ang = d_in + alm_doubles;
    cudaMemcpyAsync(d_in, alm, alm_doubles * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_ang, alpha, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_ang + nb, beta, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_ang + 2 * nb, gamma, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    const bool pinned = (size_t)nb * per * sizeof(double) < ((size_t)8 << 20) || host_is_pinned(out);
    // at least 1024 rotations per slab: the GEMMs need >= ~300 column tiles to fill the GPU (4.3 GB per slab at lmax = 512)
    long long slab = std::min<long long>(nb, std::max<long long>(1024, (long long)(h->slab_bytes / (per * sizeof(double)))));
    const int nbuf = slab < nb ? 2 : 1;
    auto grow = [](double **buf, size_t *have, size_t need) -> cudaError_t {
        if (*have >= need) return cudaSuccess;
        if (*buf) cudaFree(*buf);
        *buf = nullptr; *have = 0;
        cudaError_t e = cudaMalloc(buf, need);
        if (e == cudaSuccess) *have = need;
        return e;
    };
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
        if (k >= nbuf) cudaStreamWaitEvent(h->stream, h->ev_copied[i], 0);
        rc = rotate_device(h, lmax, d_in + (alm_stride ? 2 * (size_t)alm_stride * b0 : 0), alm_stride, d_ang + b0, d_ang + nb + b0,
                           d_ang + 2 * nb + b0, cnt, h->ws_slab[i]);
        if (rc) break;
        cudaEventRecord(h->ev_done[i], h->stream);
        cudaStreamWaitEvent(h->copy_stream, h->ev_done[i], 0);
        rc = copy_out(h, (char *)(out + (size_t)b0 * per), (const char *)h->ws_slab[i], (size_t)cnt * per * sizeof(double), pinned);
        cudaEventRecord(h->ev_copied[i], h->copy_stream);
    }
    cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->copy_stream);
    cudaFree(d_in);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(WD_ERR_CUDA, "rotation failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return WD_OK;
}

int wd_set_slab_bytes(wd_handle_t h, int64_t bytes)
{
    if (!h) return fail(WD_ERR_ASSERT, "null handle");
    if (bytes < (1 << 20)) return fail(WD_ERR_ASSERT, "slab size must be at least 1 MiB");
    if (!h->children.empty()) return for_each_child(h, [bytes](wd_handle_t c, int) { return wd_set_slab_bytes(c, bytes); });
    std::lock_guard<std::mutex> lk(h->mtx);
    h->slab_bytes = (size_t)bytes;
    return WD_OK;
}

int wd_gather(wd_handle_t h, const void *const *src, const int64_t *bytes, int dst_child, void *dst)
{
    // optional final gather of per-device output slabs onto one device (peer copies over NVLink, not part of the hot path)
    if (!h || !src || !bytes || !dst) return fail(WD_ERR_ASSERT, "null argument");
    const int R = wd_multi_size(h);
    if (dst_child < 0 || dst_child >= R) return fail(WD_ERR_ASSERT, "destination child %d out of range", dst_child);
    wd_handle_t d = h->children.empty() ? h : h->children[dst_child];
    std::lock_guard<std::mutex> lk(d->mtx);
    CU(cudaSetDevice(d->device));
    size_t off = 0;
    for (int r = 0; r < R; ++r) {
        wd_handle_t c = h->children.empty() ? h : h->children[r];
        if (bytes[r] < 0) return fail(WD_ERR_ASSERT, "negative slab size");
        if (bytes[r] > 0) CU(cudaMemcpyPeerAsync((char *)dst + off, d->device, src[r], c->device, (size_t)bytes[r], d->stream));
        off += (size_t)bytes[r];
    }
    CU(cudaStreamSynchronize(d->stream));
    return WD_OK;
}

// ---- persisted eigenvector tables: flat file, one record per 2j -------------------------------
namespace {
struct TabHeader { char magic[8]; int32_t version, count; };
struct TabRecord { int32_t twoj, Q, ldu, Kp; };
const char kTabMagic[8] = {'W', 'D', 'B', '2', '0', '0', 'T', 'B'};
}

int wd_tables_save(wd_handle_t h, const char *path)
{
    if (!h || !path) return fail(WD_ERR_ASSERT, "null argument");
    if (!h->children.empty()) h = h->children[0];
    std::lock_guard<std::mutex> lk(h->mtx);
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    FILE *f = fopen(path, "wb");
    if (!f) return fail(WD_ERR_ARGUMENT, "cannot open %s for writing", path);
    TabHeader hd;
    memcpy(hd.magic, kTabMagic, 8);
    hd.version = 1; hd.count = (int32_t)h->plans.size();
    bool ok = fwrite(&hd, sizeof hd, 1, f) == 1;
    std::vector<double> buf;
    for (auto &kv : h->plans) {
        const PlanDev &d = kv.second.dev;
        TabRecord r{d.twoj, d.Q, d.ldu, d.Kp};
        const size_t n = (size_t)d.Q * d.ldu + 2 * (size_t)d.Kp;       // uq, mu, w are one allocation
        buf.resize(n);
        if (cudaMemcpy(buf.data(), kv.second.uq, n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; break; }
        ok = ok && fwrite(&r, sizeof r, 1, f) == 1 && fwrite(buf.data(), sizeof(double), n, f) == n;
        if (!ok) break;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(WD_ERR_CUDA, "writing %s failed", path);
    return WD_OK;
}

static int tables_load_single(wd_handle_t h, const char *path)
{
    std::lock_guard<std::mutex> lk(h->mtx);
    CU(cudaSetDevice(h->device));
    FILE *f = fopen(path, "rb");
    if (!f) return fail(WD_ERR_ARGUMENT, "cannot open %s", path);
    TabHeader hd;
    if (fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, kTabMagic, 8) != 0 || hd.version != 1 || hd.count < 0) {
        fclose(f);
        return fail(WD_ERR_ARGUMENT, "%s is not a table file of this library (version 1)", path);
    }
    // the records that are not cached yet go to ONE host buffer, one device block, one copy
    std::vector<double> host;
    std::vector<TabRecord> recs;
    std::vector<size_t> offs;
    int rc = WD_OK;
    for (int i = 0; i < hd.count && rc == WD_OK; ++i) {
        TabRecord r;
        if (fread(&r, sizeof r, 1, f) != 1) { rc = fail(WD_ERR_ARGUMENT, "%s is truncated", path); break; }
        if (r.twoj < 0 || r.twoj > kMaxTwoJ) { rc = fail(WD_ERR_ARGUMENT, "%s: bad 2j = %d", path, r.twoj); break; }
        PlanDev d{};
        plan_dims(r.twoj, d);
        if (r.Q != d.Q || r.ldu != d.ldu || r.Kp != d.Kp) { rc = fail(WD_ERR_ARGUMENT, "%s: layout of 2j = %d does not match this build", path, r.twoj); break; }
        const size_t n = (size_t)r.Q * r.ldu + 2 * (size_t)r.Kp;
        if (h->plans.count(r.twoj)) {                                 // already cached: keep it, skip the record
            if (fseek(f, (long)(n * sizeof(double)), SEEK_CUR) != 0) { rc = fail(WD_ERR_ARGUMENT, "%s is truncated", path); break; }
            continue;
        }
        const size_t off = host.size();
        host.resize(off + ((n + 31) & ~(size_t)31), 0.0);
        if (fread(host.data() + off, sizeof(double), n, f) != n) { rc = fail(WD_ERR_ARGUMENT, "%s is truncated", path); break; }
        recs.push_back(r); offs.push_back(off);
    }
    fclose(f);
    if (rc || recs.empty()) return rc;
    double *block = nullptr;
    if (cudaMalloc(&block, host.size() * sizeof(double)) != cudaSuccess) return fail(WD_ERR_NOMEM, "cudaMalloc of %zu bytes of tables failed", host.size() * sizeof(double));
    if (cudaMemcpy(block, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(block); return fail(WD_ERR_CUDA, "upload of the tables failed"); }
    h->table_blocks.push_back(block);
    for (size_t i = 0; i < recs.size(); ++i) {
        Plan pl;
        plan_dims(recs[i].twoj, pl.dev);
        const size_t nuq = (size_t)recs[i].Q * recs[i].ldu;
        pl.uq = block + offs[i]; pl.mu = pl.uq + nuq; pl.w = pl.mu + recs[i].Kp;
        pl.dev.uq = pl.uq; pl.dev.mu = pl.mu; pl.dev.w = pl.w;
        h->plans[recs[i].twoj] = pl;
    }
    h->table_dirty = true;
    return WD_OK;
}

int wd_tables_load(wd_handle_t h, const char *path)
{
    if (!h || !path) return fail(WD_ERR_ASSERT, "null argument");
    if (!h->children.empty()) return for_each_child(h, [path](wd_handle_t c, int) { return tables_load_single(c, path); });
    return tables_load_single(h, path);
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
    case 3: return k2c_shape_ok(twoj) ? (int64_t)k2c_smem_layout(p.Q, p.Kp, p.ldu, p.N).total * (int64_t)sizeof(double) : ((int64_t)1 << 40);
    case 4: return k3t_shape_ok(twoj) ? (int64_t)k3t_smem_layout(p.Q, p.Kp, p.ldu, kK3TWarps).total * (int64_t)sizeof(double) : ((int64_t)1 << 40);
    default: return -1;
    }
}

int wd_fp64_peak(wd_handle_t h, int kind, int iters, double *tflops)
{
    if (!h || !tflops) return fail(WD_ERR_ASSERT, "null argument");
    if (!h->children.empty()) h = h->children[0];
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
    // kind 0: DMMA.8x8x4 (512 flops per warp instruction), 1: DFMA, 2 / 3: 1e12 sin evaluations / rotation steps per second
    const double flops = (kind == 0) ? warps * iters * 8.0 * 512.0 : (kind == 1) ? (double)blocks * threads * iters * 16.0 * 2.0
                                                                                 : (double)blocks * threads * iters;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return WD_OK;
}

}  // extern "C"
