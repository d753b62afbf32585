/* Signal preparation on the host (stays on the CPU by design, as in the reference).
 *
 * Same results as src/util.c:92-204 (quantilef, medianf, madf, medmad_normalise_array)
 * and src/scrappie_common.c:5-73 (trim_raw_by_mad, trim_and_segment_raw): quantiles by
 * linear interpolation on a sorted copy, MAD scaled by 1.4826, leading / trailing chunks
 * dropped while their MAD does not exceed the requested quantile of chunk MADs.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "scrappie_b200.h"

static int float_order(const void *a, const void *b) {
    const float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

void quantilef(const float *x, size_t nx, float *p, size_t np) {
    if (NULL == p) return;
    float *sorted = (NULL != x && nx > 0) ? malloc(nx * sizeof(float)) : NULL;
    if (NULL == sorted) {
        for (size_t i = 0; i < np; i++) p[i] = NAN;
        return;
    }
    memcpy(sorted, x, nx * sizeof(float));
    qsort(sorted, nx, sizeof(float), float_order);
    for (size_t i = 0; i < np; i++) {
        /* position p*(nx-1) is evaluated in float; the blend in double, except the
         * product frac*hi which is a float product -- all as in src/util.c:121-128 */
        const float where = p[i] * (nx - 1);
        const size_t lo = (size_t)where;
        const float frac = where - lo;
        if (lo < nx - 1) {
            const float upper = frac * sorted[lo + 1];
            p[i] = (float)((1.0 - frac) * sorted[lo] + upper);
        } else {
            p[i] = sorted[lo];
        }
    }
    free(sorted);
}

float medianf(const float *x, size_t n) {
    float q = 0.5f;
    quantilef(x, n, &q, 1);
    return q;
}

float madf(const float *x, size_t n, const float *med) {
    const float mad_scale = 1.4826;
    if (NULL == x) return NAN;
    if (1 == n) return 0.0f;
    float *dev = malloc(n * sizeof(float));
    if (NULL == dev) return NAN;
    const float centre = (NULL == med) ? medianf(x, n) : *med;
    for (size_t i = 0; i < n; i++) dev[i] = fabsf(x[i] - centre);
    const float mad = medianf(dev, n);
    free(dev);
    return mad * mad_scale;
}

void medmad_normalise_array(float *x, size_t n) {
    if (NULL == x) return;
    if (1 == n) {
        x[0] = 0.0f;
        return;
    }
    const float med = medianf(x, n);
    const float mad = madf(x, n, &med);
    for (size_t i = 0; i < n; i++) x[i] = (x[i] - med) / mad;
}

raw_table trim_raw_by_mad(raw_table rt, size_t chunk_size, float perc) {
    const raw_table failed = {0};
    if (NULL == rt.raw || chunk_size < 2) return failed;
    const size_t nchunk = (rt.end - rt.start) / chunk_size;
    rt.end = nchunk * chunk_size;               /* sic: not offset by start (scrappie_common.c:46) */
    if (0 == nchunk) return rt;

    float *mads = malloc(nchunk * sizeof(float));
    if (NULL == mads) return failed;
    for (size_t c = 0; c < nchunk; c++) mads[c] = madf(rt.raw + rt.start + c * chunk_size, chunk_size, NULL);
    float thresh = perc;
    quantilef(mads, nchunk, &thresh, 1);
    for (size_t c = 0; c < nchunk && !(mads[c] > thresh); c++) rt.start += chunk_size;
    for (size_t c = nchunk; c > 0 && !(mads[c - 1] > thresh); c--) rt.end -= chunk_size;
    free(mads);
    return rt;
}

raw_table trim_and_segment_raw(raw_table rt, size_t trim_start, size_t trim_end,
                               size_t varseg_chunk, float varseg_thresh) {
    const raw_table failed = {0};
    if (NULL == rt.raw) return failed;
    rt = trim_raw_by_mad(rt, varseg_chunk, varseg_thresh);
    if (NULL == rt.raw) return failed;
    rt.start = (rt.n - rt.start) > trim_start ? rt.start + trim_start : rt.n;
    rt.end = (rt.end > trim_end) ? rt.end - trim_end : 0;
    if (rt.start >= rt.end) {
        free(rt.raw);                           /* ownership rule of the reference (:14-17) */
        return failed;
    }
    return rt;
}
