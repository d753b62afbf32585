/* TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * Times the reference's own CPU implementation (oracle/_ref/libscrappie_ref.so, built
 * from /root/reference sources) over a set of reads, the way `scrappie raw` runs them:
 * one read per OpenMP thread, `parallel for schedule(dynamic)` (src/scrappie_raw.c:355-387),
 * single-threaded BLAS (README.md:66-71), and per read the sequence of calculate_post
 * (src/scrappie_raw.c:265-315) on an already trimmed + normalised signal:
 *   posterior -> decode_transducer -> homopolymer_path(mean) -> overlapper       (rgrgr_*)
 *   transitions -> decode_crf -> crfpath_to_basecall                             (rnnrf_r94)
 * Only prototypes are declared here; no reference source is included or copied.
 */
#include <math.h>
#include <omp.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { size_t nr, nrq, nc, stride; float *f; } ref_mat;
typedef struct { char *uuid; size_t n, start, end; float *raw; } ref_raw_table;
typedef ref_mat *(*ref_post_fn)(const ref_raw_table, float, float, float, bool);

int get_raw_model(const char *);
ref_post_fn get_posterior_function(int);
ref_mat *free_scrappie_matrix(ref_mat *);
float decode_transducer(const ref_mat *, float, float, float, int *, bool);
float decode_crf(const ref_mat *, int *);
int homopolymer_path(const ref_mat *, int *, int);
char *overlapper(const int *, size_t, int, int *);
char *crfpath_to_basecall(const int *, size_t, int *);
void scipy_openblas_set_num_threads(int);
char *scipy_openblas_get_config(void);

static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Returns wall seconds; fills total bases / blocks called.  bases_out (optional): per-read
 * malloc'd strings the caller frees with ref_bench_free. */
double ref_bench_run(const char *model, const float *concat, const size_t *offset, const size_t *nsample,
                     size_t nread, int nthreads, size_t *nbase_total, size_t *nblock_total, char **bases_out) {
    const int mt = get_raw_model(model);
    ref_post_fn post_fn = get_posterior_function(mt);
    const bool crf = (0 == strcmp(model, "rnnrf_r94"));
    scipy_openblas_set_num_threads(1);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    size_t nbase = 0, nblk = 0;
    const double t0 = now();
#pragma omp parallel for schedule(dynamic) reduction(+ : nbase, nblk)
    for (size_t r = 0; r < nread; r++) {
        ref_raw_table rt = {NULL, nsample[r], 0, nsample[r], (float *)(concat + offset[r])};
        ref_mat *post = post_fn(rt, 1e-5f, 1.0f, 1.0f, true);
        if (NULL == post) continue;
        const size_t nb = post->nc;
        int *path = calloc(nb + 1, sizeof(int));
        int *pos = calloc(nb + 1, sizeof(int));
        char *bases;
        if (!crf) {
            (void)decode_transducer(post, 0.0f, 0.0f, 2.0f, path, false);
            (void)homopolymer_path(post, path, 1);
            bases = overlapper(path, nb + 1, (int)post->nr - 1, pos);
        } else {
            (void)decode_crf(post, path);
            bases = crfpath_to_basecall(path, nb, pos);
        }
        nblk += nb;
        if (bases) nbase += strlen(bases);
        if (bases_out) bases_out[r] = bases; else free(bases);
        free(pos);
        free(path);
        free_scrappie_matrix(post);
    }
    const double dt = now() - t0;
    if (nbase_total) *nbase_total = nbase;
    if (nblock_total) *nblock_total = nblk;
    return dt;
}

void ref_bench_free(void *p) { free(p); }
/* name + version + kernel of the BLAS the reference was linked against (e.g. "OpenBLAS 0.3.31 ... SkylakeX") */
const char *ref_bench_blas_config(void) { return scipy_openblas_get_config(); }
int ref_bench_max_threads(void) { return omp_get_max_threads(); }
