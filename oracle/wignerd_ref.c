/* CPU restatement of the reference hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Line-by-line C port of jishnub/WignerD.jl v0.1.4 src/WignerD.jl:187-206
 * (_wignerdjmn: one sincos PER TERM, sequential FP64 sum), :266-295 (direct
 * wedge + three symmetry fills, in the reference's order), :335-342 (wignerD!
 * phase loop) and :245 (wignerDjmn).  Julia is not installed in this image, so
 * this port is what bench.py times as the "reference CPU path" (kind "port").
 * The eigenvectors are NOT computed here: like the reference's cache
 * (:163-176) they are produced once per j by LAPACK zheevr (through
 * oracle/wignerd_oracle.py: scipy.linalg.eigh(driver="evr")) and handed in as
 * the StructArray the reference stores: vre/vim[mu_idx + N*m_idx] = <m|mu>.
 *
 * Never linked into, imported or called by the product path.
 *
 * Build: gcc -O2 -fopenmp -fPIC -shared -o oracle/libwignerd_ref.so oracle/wignerd_ref.c -lm
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* src/WignerD.jl:187-206 */
static double elem(int twoj, int mind, int nind, double beta,
                   const double *vre, const double *vim)
{
    const int N = twoj + 1;
    const double *rm = vre + (size_t)N * mind, *im = vim + (size_t)N * mind;
    const double *rn = vre + (size_t)N * nind, *in = vim + (size_t)N * nind;
    double acc = 0.0;
    for (int k = 0; k < N; ++k) {
        double mu = 0.5 * (double)(2 * k - twoj);
        double s, c;
        sincos(-mu * beta, &s, &c);                        /* :199 */
        double T1 = rm[k] * rn[k] + im[k] * in[k];         /* :200 */
        double T2 = im[k] * rn[k] - rm[k] * in[k];         /* :201 */
        acc += c * T1 - s * T2;                            /* :202 */
    }
    return acc;
}

/* src/WignerD.jl:266-295, generic angle.  d is N x N column-major, d[mi + N*ni]. */
static void matrix(int twoj, double beta, const double *vre, const double *vim, double *d)
{
    const int N = twoj + 1;
#define IDX(t) (((t) + twoj) / 2)
#define D(tm, tn) d[IDX(tm) + (size_t)N * IDX(tn)]
#define SGN(tm, tn) (((((tm) - (tn)) / 2) & 1) ? -1.0 : 1.0)
    for (int tn = -twoj; tn <= 0; tn += 2)                 /* :271-273 */
        for (int tm = tn; tm <= -tn; tm += 2)
            D(tm, tn) = elem(twoj, IDX(tm), IDX(tn), beta, vre, vim);
    const int start = (twoj % 2 == 0) ? 2 : 1;
    for (int tn = start; tn <= twoj; tn += 2)              /* :277-281 */
        for (int tm = -tn; tm <= tn; tm += 2)
            D(tm, tn) = SGN(tm, tn) * D(-tm, -tn);
    for (int tm = start; tm <= twoj; tm += 2)              /* :284-288 */
        for (int tn = -tm; tn <= tm; tn += 2)
            D(tm, tn) = SGN(tm, tn) * D(tn, tm);
    for (int tm = -twoj; tm <= -2; tm += 2)                /* :290-292 */
        for (int tn = tm; tn <= -tm; tn += 2)
            D(tm, tn) = D(-tn, -tm);
#undef IDX
#undef D
#undef SGN
}

int wdref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* wignerd(j, beta) for nb angles; out = nb matrices, each N*N column-major (zeros-initialised like :321). */
void wdref_wignerd_batch(int twoj, const double *vre, const double *vim,
                         const double *beta, int64_t nb, double *out, int nthreads)
{
    const size_t NN = (size_t)(twoj + 1) * (twoj + 1);
#ifdef _OPENMP
    omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < nb; ++b) {
        double *d = out + NN * (size_t)b;
        memset(d, 0, NN * sizeof(double));
        matrix(twoj, beta[b], vre, vim, d);
    }
}

/* wignerD(j, alpha, beta, gamma): out = nb matrices of N*N interleaved (re, im), column-major. :335-342 */
void wdref_wignerD_batch(int twoj, const double *vre, const double *vim,
                         const double *alpha, const double *beta, const double *gamma,
                         int64_t nb, double *out, double *scratch_unused, int nthreads)
{
    (void)scratch_unused;
    const int N = twoj + 1;
    const size_t NN = (size_t)N * N;
#ifdef _OPENMP
    omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
#endif
#pragma omp parallel
    {
        double *d = (double *)__builtin_malloc(NN * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < nb; ++b) {
            memset(d, 0, NN * sizeof(double));
            matrix(twoj, beta[b], vre, vim, d);
            double *D = out + 2 * NN * (size_t)b;
            for (int ni = 0; ni < N; ++ni) {
                double n = 0.5 * (double)(2 * ni - twoj);
                for (int mi = 0; mi < N; ++mi) {
                    double m = 0.5 * (double)(2 * mi - twoj);
                    double s, c;
                    sincos(-(m * alpha[b] + n * gamma[b]), &s, &c);     /* :339 cis */
                    double v = d[mi + (size_t)N * ni];
                    D[2 * (mi + (size_t)N * ni)] = v * c;
                    D[2 * (mi + (size_t)N * ni) + 1] = v * s;
                }
            }
        }
        __builtin_free(d);
    }
}

/* WignerD.wignerdjmn over tuples (:214-216); vre_by_twoj[twoj] / vim_by_twoj[twoj] = cached eigenvectors. */
void wdref_wignerdjmn_batch(const int32_t *twoj, const int32_t *twom, const int32_t *twon,
                            const double *beta, int64_t n,
                            const double *const *vre_by_twoj, const double *const *vim_by_twoj,
                            double *out, int nthreads)
{
#ifdef _OPENMP
    omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n; ++t) {
        int tj = twoj[t];
        out[t] = elem(tj, (twom[t] + tj) / 2, (twon[t] + tj) / 2, beta[t],
                      vre_by_twoj[tj], vim_by_twoj[tj]);
    }
}
