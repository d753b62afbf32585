/* A C caller of the batch drop-in call, the way `scrappie raw` drives its reads: a team of host threads, each taking the
 * next batch of reads and calling sb2_basecall_batch on it (the reference's loop is `#pragma omp parallel for
 * schedule(dynamic)` over files, src/scrappie_raw.c:355-387, with one calculate_post per read).
 *
 * Built as scrappie_b200/libsb2_caller.so by scrappie_b200/csrc/Makefile; bench.py times it for the `e2e` figure (no
 * Python inside the timed region) and tests/ check its results against per-read calls.  Only the public C ABI of
 * include/scrappie_b200.h is used.
 *
 *   sb2_caller_run(engine, model, signals, nsample, batch_start, nbatch, order, nstep, nthread, params,
 *                  bases_out, score_out, nbase_total)
 *     signals / nsample   every read of the workload (ordinary host memory)
 *     batch_start         nbatch + 1 offsets: batch k = reads [batch_start[k], batch_start[k + 1])
 *     order               the order in which the batches of a step are handed out (longest first), or NULL
 *     nstep               how many times the whole workload is called
 *     bases_out           (optional) one malloc'd base string per read from the LAST step; caller frees each
 *   returns wall-clock seconds of all nstep passes, or a negative number on failure.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "scrappie_b200.h"

typedef struct {
    sb2_engine *eng;
    enum raw_model_type model;
    const float *const *signals;
    const size_t *nsample;
    const size_t *batch_start;
    const int *order;
    int nbatch, nstep;
    const sb2_params *params;
    char **bases_out;
    float *score_out;
    atomic_long next;               /* next work item: step * nbatch + position in `order` */
    atomic_long nbase, failed;
    char errmsg[512];               /* sb2_last_error() of the first call that failed (the library's message is per thread) */
} caller_job;

static char caller_error[512];
const char *sb2_caller_last_error(void) { return caller_error; }

static void caller_fail(caller_job *job, const char *what) {
    if (0 == atomic_fetch_add(&job->failed, 1)) {
        const char *msg = sb2_last_error();
        snprintf(job->errmsg, sizeof(job->errmsg), "%s: %s", what, (msg && msg[0]) ? msg : "(no message)");
    }
}

static void *caller_worker(void *arg) {
    caller_job *job = arg;
    const long nitem = (long)job->nbatch * job->nstep;
    size_t cap = 0;
    sb2_call *calls = NULL;
    for (;;) {
        const long item = atomic_fetch_add(&job->next, 1);
        if (item >= nitem) break;
        const int step = (int)(item / job->nbatch), pos = (int)(item % job->nbatch);
        const int k = job->order ? job->order[pos] : pos;
        const size_t r0 = job->batch_start[k], n = job->batch_start[k + 1] - r0;
        if (n > cap) {
            free(calls);
            calls = malloc(n * sizeof(sb2_call));
            cap = n;
            if (NULL == calls) { caller_fail(job, "out of host memory"); break; }
        }
        const int ncalled = sb2_basecall_batch(job->eng, job->model, job->signals + r0, job->nsample + r0, n, job->params, calls);
        if (ncalled < 0) { caller_fail(job, "sb2_basecall_batch"); break; }
        long nb = 0;
        const int keep = (step == job->nstep - 1);
        for (size_t i = 0; i < n; i++) {
            nb += (long)calls[i].nbase;
            if (keep && job->score_out) job->score_out[r0 + i] = calls[i].score;
            if (keep && job->bases_out) job->bases_out[r0 + i] = calls[i].bases;    /* ownership passes to the caller */
            else free(calls[i].bases);
        }
        if (keep) atomic_fetch_add(&job->nbase, nb);
    }
    free(calls);
    return NULL;
}

double sb2_caller_run(sb2_engine *eng, int model, const float *const *signals, const size_t *nsample,
                      const size_t *batch_start, int nbatch, const int *order, int nstep, int nthread,
                      const sb2_params *params, char **bases_out, float *score_out, size_t *nbase_total) {
    if (NULL == eng || NULL == signals || NULL == nsample || NULL == batch_start || nbatch <= 0 || nstep <= 0 ||
        nthread <= 0 || NULL == params)
        return -1.0;
    caller_job job = {eng, (enum raw_model_type)model, signals, nsample, batch_start, order, nbatch, nstep, params,
                      bases_out, score_out};
    job.errmsg[0] = '\0';
    atomic_init(&job.next, 0);
    atomic_init(&job.nbase, 0);
    atomic_init(&job.failed, 0);
    if (nthread > 256) nthread = 256;
    pthread_t th[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int started = 0;
    for (; started < nthread; started++)
        if (0 != pthread_create(&th[started], NULL, caller_worker, &job)) break;
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (nbase_total) *nbase_total = (size_t)atomic_load(&job.nbase);
    if (0 == started || atomic_load(&job.failed) > 0) {
        snprintf(caller_error, sizeof(caller_error), "%s", 0 == started ? "could not start a host thread" : job.errmsg);
        return -1.0;
    }
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void sb2_caller_free(void *p) { free(p); }
