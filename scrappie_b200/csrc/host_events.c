/* Event features on the host (stay on the CPU as in the reference: 4 floats per event).
 *
 * nanonet_features_from_events + studentise_features_kahan, src/nnfeatures.c:47-115: per event
 * (mean, stdv, length, |mean - next mean|), studentised with Kahan-compensated sums, the variance
 * inverted with the SSE reciprocal-square-root approximation exactly as the reference does (so the
 * result is the reference's on the same CPU).  Plain -O2 without -march: no FMA contraction.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <xmmintrin.h>

#include "scrappie_b200.h"

scrappie_matrix nanonet_features_from_events(const event_table et, bool normalise) {
    if (NULL == et.event || et.end <= et.start) return NULL;
    const size_t nevent = et.end - et.start, offset = et.start;
    scrappie_matrix features = make_scrappie_matrix(4, nevent);
    if (NULL == features) return NULL;
    __m128 *v = (__m128 *)features->data.f;
    for (size_t ev = 0; ev + 1 < nevent; ev++)
        v[ev] = _mm_setr_ps(et.event[ev + offset].mean, et.event[ev + offset].stdv, et.event[ev + offset].length,
                            fabs(et.event[ev + offset].mean - et.event[ev + offset + 1].mean));
    v[nevent - 1] = _mm_setr_ps(et.event[et.end - 1].mean, et.event[et.end - 1].stdv, et.event[et.end - 1].length, 0.0f);
    if (!normalise) return features;

    __m128 sum = _mm_setzero_ps(), sumsq = _mm_setzero_ps(), comp = _mm_setzero_ps(), compsq = _mm_setzero_ps();
    for (size_t ev = 0; ev < nevent; ev++) {
        const __m128 d1 = _mm_sub_ps(v[ev], comp);
        const __m128 sum_tmp = _mm_add_ps(sum, d1);
        comp = _mm_sub_ps(_mm_sub_ps(sum_tmp, sum), d1);
        sum = sum_tmp;
        const __m128 d2 = _mm_sub_ps(_mm_mul_ps(v[ev], v[ev]), compsq);
        const __m128 sumsq_tmp = _mm_add_ps(sumsq, d2);
        compsq = _mm_sub_ps(_mm_sub_ps(sumsq_tmp, sumsq), d2);
        sumsq = sumsq_tmp;
    }
    const __m128 nf = _mm_set1_ps((float)(int)nevent);
    sum = _mm_div_ps(sum, nf);
    sumsq = _mm_div_ps(sumsq, nf);
    sumsq = _mm_sub_ps(sumsq, _mm_mul_ps(sum, sum));
    sumsq = _mm_rsqrt_ps(sumsq);
    sum = _mm_mul_ps(sum, sumsq);
    for (size_t ev = 0; ev < nevent; ev++) v[ev] = _mm_sub_ps(_mm_mul_ps(sumsq, v[ev]), sum);
    return features;
}

/* ------------------------------------------------------------------------------------------------
 * detect_events (src/event_detection.c:24-320): segmentation of the raw signal into events, the input of
 * nanonet_posterior.  Host code as in the reference.  Same arithmetic, own structure:
 *   1. prefix sums of x and x^2 in double;
 *   2. two windowed Welch t-statistics (window 3 and 6 by default) in float, zero within a window of either end;
 *   3. one left-to-right pass of a pair of peak trackers -- the short-window tracker masks the long one while it is
 *      about to fire -- emitting a boundary once a peak above its threshold has dropped by `peak_height` and is more
 *      than half a window old;
 *   4. events between consecutive boundaries with mean / stdv from the prefix sums.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const float *tstat;
    float threshold;
    size_t window;
    size_t masked_to;
    long peak_pos;              /* -1: no candidate peak yet */
    float peak_value;
    bool confirmed;
} sb2_peak_tracker;

static float *windowed_tstat(const double *sum, const double *sumsq, size_t n, size_t w) {
    float *t = calloc(n ? n : 1, sizeof(float));
    if (NULL == t || n < 2 * w || w < 2) return t;      /* all zero: the statistic is undefined */
    const float wf = (float)w;
    for (size_t i = w; i <= n - w; i++) {
        double s1 = sum[i], q1 = sumsq[i];
        if (i > w) { s1 -= sum[i - w]; q1 -= sumsq[i - w]; }
        const float s2 = (float)(sum[i + w] - sum[i]);
        const float q2 = (float)(sumsq[i + w] - sumsq[i]);
        const float m1 = s1 / wf, m2 = s2 / wf;
        float var = q1 / wf - m1 * m1 + q2 / wf - m2 * m2;
        var = fmaxf(var, FLT_MIN);
        const float dm = m2 - m1;
        t[i] = fabs(dm) / sqrt(var / wf);
    }
    /* the reference zeroes a window at both ends AFTER nothing was written there, except position n - w, which the
     * main loop reaches: it is computed, not zero (its loop runs to i <= n - w while the fudge covers n - w .. n - 1
     * BEFORE the loop).  Order matters: zero first, then compute -- done above by calloc + the loop. */
    return t;
}

static void tracker_reset(sb2_peak_tracker *d) {
    d->peak_pos = -1;
    d->peak_value = FLT_MAX;
    d->confirmed = false;
}

event_table detect_events(raw_table const rt, detector_param const p) {
    event_table et = {0, 0, 0, NULL};
    if (NULL == rt.raw || rt.end <= rt.start) return et;
    const size_t n = rt.end - rt.start;
    const float *x = rt.raw + rt.start;
    double *sum = calloc(n + 1, sizeof(double)), *sumsq = calloc(n + 1, sizeof(double));
    size_t *bounds = calloc(n, sizeof(size_t));
    float *t1 = NULL, *t2 = NULL;
    if (NULL == sum || NULL == sumsq || NULL == bounds) goto done;
    for (size_t i = 0; i < n; i++) {
        sum[i + 1] = sum[i] + x[i];
        sumsq[i + 1] = sumsq[i] + x[i] * x[i];          /* float product, double accumulation, as the reference */
    }
    t1 = windowed_tstat(sum, sumsq, n, p.window_length1);
    t2 = windowed_tstat(sum, sumsq, n, p.window_length2);
    if (NULL == t1 || NULL == t2) goto done;

    sb2_peak_tracker trk[2] = {{t1, p.threshold1, p.window_length1, 0, -1, FLT_MAX, false},
                               {t2, p.threshold2, p.window_length2, 0, -1, FLT_MAX, false}};
    size_t nbound = 0;
    for (size_t i = 0; i < n; i++) {
        for (int k = 0; k < 2; k++) {
            sb2_peak_tracker *d = &trk[k];
            if (d->masked_to >= i) continue;
            const float v = d->tstat[i];
            if (d->peak_pos < 0) {
                if (v < d->peak_value) d->peak_value = v;                               /* still descending */
                else if (v - d->peak_value > p.peak_height) { d->peak_value = v; d->peak_pos = (long)i; }
                continue;
            }
            if (v > d->peak_value) { d->peak_value = v; d->peak_pos = (long)i; }
            if (0 == k && d->peak_value > d->threshold) {                                /* short window dominates */
                trk[1].masked_to = (size_t)d->peak_pos + d->window;
                tracker_reset(&trk[1]);
            }
            if (d->peak_value - v > p.peak_height && d->peak_value > d->threshold) d->confirmed = true;
            if (d->confirmed && (i - (size_t)d->peak_pos) > d->window / 2) {
                bounds[nbound++] = (size_t)d->peak_pos;
                d->peak_pos = -1;
                d->peak_value = v;
                d->confirmed = false;
            }
        }
    }

    /* the reference counts boundaries by scanning the zero-padded array for entries in (0, n) */
    size_t nev = 1;
    for (size_t i = 0; i < n; i++) if (bounds[i] > 0 && bounds[i] < n) nev++;
    et.event = calloc(nev, sizeof(event_t));
    if (NULL == et.event) goto done;
    et.n = nev;
    et.start = 0;
    et.end = nev;
    for (size_t e = 0; e < nev; e++) {
        /* first event starts at 0, last ends at n.  (Without any boundary the reference indexes peaks[n - 2] with
         * n == 1, i.e. out of bounds; here that case is one event covering the whole signal.) */
        const size_t lo = (0 == e) ? 0 : bounds[e - 1];
        const size_t hi = (e + 1 == nev) ? n : bounds[e];
        event_t *ev = &et.event[e];
        ev->start = (uint64_t)lo;
        ev->length = (float)(hi - lo);
        ev->mean = (float)(sum[hi] - sum[lo]) / ev->length;
        const float dsq = (float)(sumsq[hi] - sumsq[lo]);
        const float var = dsq / ev->length - ev->mean * ev->mean;
        ev->stdv = sqrtf(fmaxf(var, 0.0f));
        ev->pos = -1;
        ev->state = -1;
    }
done:
    free(t2); free(t1); free(bounds); free(sumsq); free(sum);
    return et;
}
