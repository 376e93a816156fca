/* wignerd_b200.h -- C ABI of libwignerd_b200.so, the B200 (sm_100a) Wigner-d/D engine.
 *
 * This is the drop-in boundary for the hot path of jishnub/WignerD.jl v0.1.4.  The
 * reference has no FFI of its own; its boundary is the Julia function API
 * (src/WignerD.jl:9-12 exports, README.md:24,33 qualified names).  Every entry
 * point below names the reference function it replaces; julia/WignerD.jl shows the
 * `ccall` binding, wignerd/jl_b200/api.py the ctypes binding used by the tests.
 *
 * Conventions (identical to the reference):
 *   - j is passed as twoj = 2j (int32), so integer and half-integer j share one
 *     path -- the reference keys its own cache the same way, UInt(2j+1) (:164).
 *     m, n likewise as twom = 2m, twon = 2n.
 *   - a d matrix is N x N, N = 2j+1, COLUMN-MAJOR, element d[(j+m) + N*(j+n)] =
 *     d^j_{m,n}(beta), rows/cols ascending from -j (README.md:18-22).  Batches are
 *     nb such matrices back to back.  D matrices are the same with interleaved
 *     (re, im) complex128 elements.
 *   - Varshalovich phase convention, D^j_{mn} = exp(-i(m alpha + n gamma)) d^j_{mn}(beta)
 *     (src/WignerD.jl:245, :339).
 *   - no function throws or aborts across the boundary: each returns a wd_status and
 *     leaves a message for wd_last_error() (thread-local).  The Julia shim maps the
 *     codes back to the reference's exception types.
 *   - `mem` says where the caller's angle/output buffers live: WD_MEM_HOST (pageable
 *     or pinned host memory; the library stages through the device) or WD_MEM_DEVICE
 *     (device pointers on the handle's GPU; nothing is copied, work is enqueued on the
 *     handle's stream and the call returns without synchronising).
 *   - there is NO CPU fallback: every compute entry point fails with WD_ERR_CUDA when no
 *     sm_100-class device is usable.
 */
#ifndef WIGNERD_B200_H
#define WIGNERD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wd_handle_s *wd_handle_t;

typedef enum {
    WD_OK = 0,
    WD_ERR_ARGUMENT = 1, /* reference: ArgumentError (src/WignerD.jl:98,191-194) -- m/n outside -j..j */
    WD_ERR_ASSERT = 2,   /* reference: AssertionError (:96-97,148-149,267) -- j < 0, bad sizes, null ptr */
    WD_ERR_CUDA = 3,     /* CUDA runtime failure / no usable device */
    WD_ERR_NOMEM = 4,    /* device or host allocation failed */
    WD_ERR_UNSUPPORTED = 5
} wd_status;

/* beta::Real vs the special co-latitude types (src/WignerD.jl:14-46). */
typedef enum {
    WD_BETA_GENERIC = 0,
    WD_BETA_NORTH = 1,  /* NorthPole(): beta = 0,    :218-222, :247-253 */
    WD_BETA_SOUTH = 2,  /* SouthPole(): beta = pi,   :224-228, :255-264 */
    WD_BETA_EQUATOR = 3 /* Equator():   beta = pi/2, :230-237 */
} wd_beta_kind;

enum { WD_MEM_HOST = 0, WD_MEM_DEVICE = 1 };

/* ---- lifetime ------------------------------------------------------------------- */
/* One handle = one GPU + one stream + the per-j eigenvector-table cache (replaces the
 * per-Task JyEigenDict, src/WignerD.jl:137-138).  Calls on one handle are serialised by
 * an internal mutex; use one handle per thread for concurrency (mirrors the per-Task
 * cache that makes the reference safe under Threads.@threads). */
int wd_create(int device, wd_handle_t *out);
int wd_destroy(wd_handle_t h);
const char *wd_last_error(void);
int wd_abi_version(void);
/* enqueue on a caller-owned cudaStream_t (0/NULL = the handle's own stream; pass cudaStreamLegacy (0x1) for the default stream) */
int wd_set_stream(wd_handle_t h, void *cuda_stream);
int wd_synchronize(wd_handle_t h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t wd_launch_count(wd_handle_t h);
/* drop the cached per-j tables (the reference never frees its Dict) */
int wd_cache_clear(wd_handle_t h);

/* ---- multi-GPU: the (j, beta) batch shards with no inter-GPU traffic ---- */
/* One handle over ndev GPUs of this process (devices[i], or 0..ndev-1 when devices is NULL; a device may be
 * listed twice): one single-device child handle + host worker thread per entry.  wd_d_batch / wd_D_batch /
 * wd_d_multi / wd_djmn_batch / wd_Djmn_batch on such a handle take HOST buffers (mem = WD_MEM_HOST), cut the
 * angle (or tuple) batch into contiguous slabs with wd_shard_bounds (16-aligned for matrices) and run the slabs
 * concurrently; every device builds its own tables (the reference's cache is per Task the same way, :137).
 * wd_prepare / wd_cache_clear / wd_set_variant / wd_synchronize / wd_tables_load act on every child. */
int wd_create_multi(int ndev, const int *devices, wd_handle_t *out);
/* number of devices behind a handle (1 for wd_create handles) and the i-th per-device handle: callers that keep
 * their data on the devices call the batch entry points on the children with WD_MEM_DEVICE and their own slabs. */
int wd_multi_size(wd_handle_t h);
int wd_multi_child(wd_handle_t h, int i, wd_handle_t *child);
/* [lo, hi) of the slab of n batch items owned by rank of world: slab starts are multiples of align (16 = the
 * kernels' angle chunk), sizes differ by at most align, the slabs partition [0, n).  Pure host arithmetic. */
int wd_shard_bounds(int64_t n, int rank, int world, int align, int64_t *lo, int64_t *hi);
/* optional final gather (never part of the timed hot path): concatenates the per-device slabs src[r]
 * (bytes[r] bytes each, on child r's device) into dst on child dst_child's device with peer copies. */
int wd_gather(wd_handle_t h, const void *const *src, const int64_t *bytes, int dst_child, void *dst);

/* ---- spectral setup: replaces coeffi + eigen! + cache, src/WignerD.jl:145-185 ------ */
/* Build (or find cached) the tables for every 2j in the list, in ONE batched launch of
 * the tridiagonal eigensolver.  Implicitly called by everything below. */
int wd_prepare(wd_handle_t h, const int32_t *twoj_list, int32_t nj);
/* Real orthonormal eigenvectors u of J_x (J_y = S J_x S^dagger, S = diag((-i)^m)):
 * u_out is N x N column-major on the HOST, u_out[(j+m) + N*k] = u[m, mu_k], mu_k = -j+k.
 * <m|mu>_y = (-i)^m u[m,mu] up to a per-mu phase, which cancels in d.  Debug/parity. */
int wd_eigvecs(wd_handle_t h, int32_t twoj, double *u_out);

/* Persisted tables (extends the in-memory cache of Jy_eigen, :163-176): wd_tables_save writes every cached
 * table of the handle to a flat file (header + one record per 2j: dims, UQ, mu, w); wd_tables_load puts the
 * records of a file into the cache (existing entries are kept), so a cold process skips the eigensolver.
 * A file written by a different table layout is refused with WD_ERR_ARGUMENT. */
int wd_tables_save(wd_handle_t h, const char *path);
int wd_tables_load(wd_handle_t h, const char *path);

/* ---- matrices: replaces wignerd/wignerd! (:266-324) and wignerD/wignerD! (:335-357) -- */
/* nb generic angles -> nb matrices (see layout above).  beta has nb entries. */
int wd_d_batch(wd_handle_t h, int32_t twoj, const double *beta, int64_t nb, double *out, int mem);
/* out has nb * N*N complex128 (2 doubles each); in device memory it must be 16-byte aligned, as every complex128 array is
 * (WD_ERR_ASSERT otherwise). */
int wd_D_batch(wd_handle_t h, int32_t twoj, const double *alpha, const double *beta,
               const double *gamma, int64_t nb, double *out, int mem);
/* Single matrix into HOST memory with the reference's in-place semantics (wignerd!):
 * GENERIC/EQUATOR overwrite all N*N elements; NORTH writes only the diagonal and SOUTH only
 * the anti-diagonal (:247-264 -- the caller supplies zeros, as wignerd does at :321). */
int wd_d_matrix(wd_handle_t h, int32_t twoj, double beta, int beta_kind, double *d);
/* wignerD!: d as above into the complex matrix, then every element *= cis(-(m alpha + n gamma)). */
int wd_D_matrix(wd_handle_t h, int32_t twoj, double alpha, double beta, int beta_kind,
                double gamma, double *D);
/* Many j on one angle grid (spherical-harmonic rotation to lmax): for k in 0..nj-1 writes the
 * nb matrices of twoj_list[k] at out + offsets[k] (offsets in doubles; blocks may alias when the
 * caller ring-buffers).  beta: nb generic angles. */
int wd_d_multi(wd_handle_t h, const int32_t *twoj_list, int32_t nj, const double *beta, int64_t nb,
               double *out, const int64_t *offsets, int mem);

/* ---- fused consumer: rotation of spherical-harmonic coefficients (docs/src/index.md:76-84) ------ */
/* out[b][l^2 + l + m'] = sum_m D^l_{m m'}(alpha_b, beta_b, gamma_b) alm[b * alm_stride + l^2 + l + m]  for l = 0..lmax,
 * i.e. what a caller of wignerD(l, alpha, beta, gamma) for every l <= lmax does with the matrices (the reference's cfg-4
 * use case) -- computed from the eigenvectors directly (two tensor-core GEMMs per l), no D matrix is ever formed.
 * alm, out: interleaved complex128; alm_stride (in complex elements) = 0 rotates ONE shared coefficient set by all nb
 * rotations, otherwise >= (lmax+1)^2 (one set per rotation); out holds nb * (lmax+1)^2 complex numbers. */
int wd_rotate_sh(wd_handle_t h, int32_t lmax, const double *alm, int64_t alm_stride, const double *alpha,
                 const double *beta, const double *gamma, int64_t nb, double *out, int mem);

/* ---- single elements: replaces WignerD.wignerdjmn (:214-237), wignerDjmn (:245) ------ */
/* n tuples (twoj[t], twom[t], twon[t], beta[t]) -> out[t].  beta_kind applies to all tuples
 * (beta is ignored for NORTH/SOUTH/EQUATOR).  Tuples with m or n outside -j..j (or of the
 * wrong parity) make the call fail with WD_ERR_ARGUMENT (:190-193); out is then unspecified. */
int wd_djmn_batch(wd_handle_t h, const int32_t *twoj, const int32_t *twom, const int32_t *twon,
                  const double *beta, int64_t n, int beta_kind, double *out, int mem);
/* out[t] = d * cis(-(alpha m + gamma n)), interleaved complex128. */
int wd_Djmn_batch(wd_handle_t h, const int32_t *twoj, const int32_t *twom, const int32_t *twon,
                  const double *alpha, const double *beta, const double *gamma, int64_t n,
                  int beta_kind, double *out, int mem);

/* ---- introspection for tests / benches ------------------------------------------- */
/* kernel variant selection for wd_d_batch/wd_D_batch: 0 = auto (real d, table resident: angle grids that are
 * symmetric about pi/2 take the column-staged kernel with beta <-> pi - beta pairing, other grids k2_real_kernel --
 * decided on the device), 1 = force the plain FP64-ALU kernel, 2 = force the resident DMMA kernels k2_real_kernel /
 * k2_cplx_kernel, 3 = force the streaming DMMA kernel (real d), 4 = force the column-staged DMMA kernel for every grid
 * (paired when symmetric; fails if the table of j does not fit), 5 = the column-staged kernel without pairing,
 * 6 = auto with the pairing switched off (every grid is treated as generic), 7 = force the tensor-store complex-D
 * kernel k3_tma_kernel (complex D in auto mode takes it whenever 5 <= 2j <= 127 and the output is 16-byte aligned;
 * forced, a request outside its range fails with WD_ERR_UNSUPPORTED; real d is dispatched as in auto mode). */
int wd_set_variant(wd_handle_t h, int variant);
/* host path (mem = WD_MEM_HOST): bytes of output per kernel launch / device-to-host copy (default 512 MiB;
 * two such slabs are double-buffered on the device). */
int wd_set_slab_bytes(wd_handle_t h, int64_t bytes);
/* dynamic shared memory (bytes) that the DMMA kernels need for a given 2j -- kernel 0: round-1 resident real-d
 * kernel, 1: resident complex-D kernel, 2: streaming kernel (independent of j), 3: column-staged real-d kernel, 4:
 * tensor-store complex-D kernel (3, 4: a huge value when the shape of j is outside the kernel's range).  A resident kernel is used when this fits the device's opt-in
 * limit (232448 bytes on sm_100a).  Pure host arithmetic, no GPU needed; -1 on bad arguments. */
int64_t wd_k2_shared_bytes(int32_t twoj, int kernel);
/* FP64 micro-benchmarks used to calibrate the rooflines (bench.py): returns TFLOP/s of a
 * register-resident DMMA.8x8x4 (kind 0) or DFMA (kind 1) loop over the whole chip; kind 2: 1e12 full-range
 * double sin evaluations per second, kind 3: 1e12 complex rotation steps per second (the denominators of the
 * single-element kernel, cfg-5). */
int wd_fp64_peak(wd_handle_t h, int kind, int iters, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* WIGNERD_B200_H */
