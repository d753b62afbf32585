/* Event features on the host (stay on the CPU as in the reference: 4 floats per event).
 *
 * nanonet_features_from_events + studentise_features_kahan, src/nnfeatures.c:47-115: per event
 * (mean, stdv, length, |mean - next mean|), studentised with Kahan-compensated sums, the variance
 * inverted with the SSE reciprocal-square-root approximation exactly as the reference does (so the
 * result is the reference's on the same CPU).  Plain -O2 without -march: no FMA contraction.
 */
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
