// Signal preparation on the device: trim_and_segment_raw / trim_raw_by_mad (src/scrappie_common.c:5-73) and
// medmad_normalise_array (src/util.c:190-204) for a batch of reads, bit-identical to the host functions
// (csrc/host_signal.c) and hence to the reference.
//
// The reference sorts a copy of the data (qsort) to read off one or two order statistics; here the order
// statistics are SELECTED, which is exact for any data:
//   * whole-read median / MAD: one CTA per read, radix select over the order-preserving integer image of the
//     floats (4 passes of 8 bits, 256-bin shared-memory histogram) plus one pass for the next larger element;
//     the absolute deviations of the MAD are recomputed on the fly, never stored;
//   * the per-chunk MADs of the trimmer (chunks of `varseg_chunk` = 100 samples): one warp per chunk with the
//     chunk in registers, bit-serial select with __reduce_add_sync (chunks of more than 256 samples take the
//     CTA path);
//   * the quantile blend mirrors src/util.c:121-128 operation by operation (float position and float product,
//     double blend), with explicit round-to-nearest intrinsics so that no FMA contraction can change a bit.
#include <math.h>

#include "kernels.h"

namespace sb2 {
namespace {

constexpr int PREP_THREADS = 256;
constexpr int WARP_MAXE = 8;                            // register-resident chunk: up to 32 * 8 samples

__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// quantile blend of src/util.c:121-128 given the two neighbouring order statistics
__device__ __forceinline__ float quantile_blend(float s0, float s1, float frac) {
    const float upper = __fmul_rn(frac, s1);
    return (float)__dadd_rn(__dmul_rn(__dsub_rn(1.0, (double)frac), (double)s0), (double)upper);
}

struct SelectScratch {
    uint32_t hist[256];
    uint32_t digit, krem, cnt_le, min_gt;
};

// k-th smallest (0-based) key of get(0..n-1), all threads of the CTA; returns the key to every thread
template <class F>
__device__ uint32_t cta_select_key(F get, int n, int k, SelectScratch &sc) {
    const int tid = threadIdx.x;
    uint32_t prefix = 0, mask = 0;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 256; i += blockDim.x) sc.hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {
            const uint32_t key = f2key(get(i));
            if ((key & mask) == prefix) atomicAdd(&sc.hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t cum = 0, kk = (uint32_t)k;
            int dgt = 255;
            for (int b = 0; b < 256; b++) {
                const uint32_t h = sc.hist[b];
                if (kk < cum + h) { dgt = b; break; }
                cum += h;
            }
            sc.digit = (uint32_t)dgt;
            sc.krem = kk - cum;
        }
        __syncthreads();
        prefix |= sc.digit << shift;
        mask |= 0xffu << shift;
        k = (int)sc.krem;
        __syncthreads();
    }
    return prefix;
}

// the order statistic after position k, given the key at k: the same key if it occurs again, else the smallest larger one
template <class F>
__device__ uint32_t cta_next_key(F get, int n, int k, uint32_t key_k, SelectScratch &sc) {
    const int tid = threadIdx.x;
    if (tid == 0) { sc.cnt_le = 0; sc.min_gt = 0xffffffffu; }
    __syncthreads();
    uint32_t cnt = 0, mn = 0xffffffffu;
    for (int i = tid; i < n; i += blockDim.x) {
        const uint32_t key = f2key(get(i));
        if (key <= key_k) cnt++;
        else mn = min(mn, key);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    mn = __reduce_min_sync(0xffffffffu, mn);
    if ((tid & 31) == 0) { atomicAdd(&sc.cnt_le, cnt); atomicMin(&sc.min_gt, mn); }
    __syncthreads();
    const uint32_t res = (sc.cnt_le >= (uint32_t)k + 2u) ? key_k : sc.min_gt;
    __syncthreads();
    return res;
}

// quantilef (src/util.c:92-133) for one probability
template <class F>
__device__ float cta_quantile(F get, int n, float p, SelectScratch &sc) {
    const float where = __fmul_rn(p, (float)(n - 1));
    const int lo = (int)where;
    const float frac = __fsub_rn(where, (float)lo);
    const uint32_t k0 = cta_select_key(get, n, lo, sc);
    if (lo < n - 1) {
        const uint32_t k1 = cta_next_key(get, n, lo, k0, sc);
        return quantile_blend(key2f(k0), key2f(k1), frac);
    }
    return key2f(k0);
}

// ---- warp versions: element e of lane l is sample l + 32 e of the chunk -----------------------------
__device__ __forceinline__ uint32_t warp_select_key(const uint32_t (&keys)[WARP_MAXE], int n, int k, int lane) {
    uint32_t prefix = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t hi_mask = (bit == 31) ? 0u : (0xffffffffu << (bit + 1));
        uint32_t c0 = 0;
#pragma unroll
        for (int e = 0; e < WARP_MAXE; e++) {
            const bool valid = lane + 32 * e < n;
            c0 += (valid && ((keys[e] & hi_mask) == prefix) && !((keys[e] >> bit) & 1u)) ? 1u : 0u;
        }
        c0 = __reduce_add_sync(0xffffffffu, c0);
        if ((uint32_t)k >= c0) { k -= (int)c0; prefix |= 1u << bit; }
    }
    return prefix;
}
__device__ __forceinline__ uint32_t warp_next_key(const uint32_t (&keys)[WARP_MAXE], int n, int k, uint32_t key_k, int lane) {
    uint32_t cnt = 0, mn = 0xffffffffu;
#pragma unroll
    for (int e = 0; e < WARP_MAXE; e++) {
        if (lane + 32 * e < n) {
            if (keys[e] <= key_k) cnt++;
            else mn = min(mn, keys[e]);
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    mn = __reduce_min_sync(0xffffffffu, mn);
    return (cnt >= (uint32_t)k + 2u) ? key_k : mn;
}
__device__ __forceinline__ float warp_median(const uint32_t (&keys)[WARP_MAXE], int n, int lane) {
    const float where = __fmul_rn(0.5f, (float)(n - 1));
    const int lo = (int)where;
    const float frac = __fsub_rn(where, (float)lo);
    const uint32_t k0 = warp_select_key(keys, n, lo, lane);
    if (lo < n - 1) return quantile_blend(key2f(k0), key2f(warp_next_key(keys, n, lo, k0, lane)), frac);
    return key2f(k0);
}
// madf(x, n, NULL) (src/util.c:160-188) of one chunk
__device__ float warp_mad(const float *x, int n, int lane) {
    if (n == 1) return 0.0f;
    float v[WARP_MAXE];
    uint32_t keys[WARP_MAXE];
#pragma unroll
    for (int e = 0; e < WARP_MAXE; e++) {
        v[e] = (lane + 32 * e < n) ? x[lane + 32 * e] : 0.0f;
        keys[e] = f2key(v[e]);
    }
    const float med = warp_median(keys, n, lane);
#pragma unroll
    for (int e = 0; e < WARP_MAXE; e++) keys[e] = f2key(fabsf(__fsub_rn(v[e], med)));
    return __fmul_rn(warp_median(keys, n, lane), 1.4826f);
}

// trim_and_segment_raw for one read per CTA; rt.start = 0 on entry as in calculate_post
__global__ void __launch_bounds__(PREP_THREADS)
trim_kernel(const float *__restrict__ raw, const int64_t *__restrict__ off, const int *__restrict__ nsample,
            int chunk, float perc, int trim_start, int trim_end, float *__restrict__ mads,
            const int64_t *__restrict__ mads_off, int2 *__restrict__ start_end) {
    __shared__ SelectScratch sc;
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *x = raw + off[r];
    const int n = nsample[r];
    const int nchunk = n / chunk;
    float *m = mads + mads_off[r];
    int start = 0, end = nchunk * chunk;
    if (nchunk > 0) {
        if (chunk <= 32 * WARP_MAXE) {
            for (int c = warp; c < nchunk; c += PREP_THREADS / 32) {
                const float mad = warp_mad(x + (size_t)c * chunk, chunk, lane);
                if (lane == 0) m[c] = mad;
            }
        } else {
            for (int c = 0; c < nchunk; c++) {
                const float *xc = x + (size_t)c * chunk;
                const float med = cta_quantile([=](int i) { return xc[i]; }, chunk, 0.5f, sc);
                const float dev = cta_quantile([=](int i) { return fabsf(__fsub_rn(xc[i], med)); }, chunk, 0.5f, sc);
                if (tid == 0) m[c] = __fmul_rn(dev, 1.4826f);
            }
        }
        __syncthreads();
        const float thresh = cta_quantile([=](int i) { return m[i]; }, nchunk, perc, sc);
        if (tid == 0) {
            for (int c = 0; c < nchunk && !(m[c] > thresh); c++) start += chunk;
            for (int c = nchunk; c > 0 && !(m[c - 1] > thresh); c--) end -= chunk;
        }
    }
    if (tid == 0) {
        start = (n - start) > trim_start ? start + trim_start : n;
        end = (end > trim_end) ? end - trim_end : 0;
        if (start >= end) { start = 0; end = 0; }       // the reference frees the read here (:14-17)
        start_end[r] = make_int2(start, end);
    }
}

// medmad_normalise_array for one read per CTA: dst[0..n) = (src[0..n) - median) / MAD
__global__ void __launch_bounds__(PREP_THREADS)
medmad_kernel(const float *__restrict__ src, const int64_t *__restrict__ src_off, float *__restrict__ dst,
              const int64_t *__restrict__ dst_off, const int *__restrict__ nsample) {
    __shared__ SelectScratch sc;
    const int r = blockIdx.x, tid = threadIdx.x;
    const float *x = src + src_off[r];
    float *y = dst + dst_off[r];
    const int n = nsample[r];
    if (n <= 0) return;
    if (n == 1) { if (tid == 0) y[0] = 0.0f; return; }
    const float med = cta_quantile([=](int i) { return x[i]; }, n, 0.5f, sc);
    const float mad = __fmul_rn(cta_quantile([=](int i) { return fabsf(__fsub_rn(x[i], med)); }, n, 0.5f, sc), 1.4826f);
    for (int i = tid; i < n; i += blockDim.x) y[i] = __fdiv_rn(__fsub_rn(x[i], med), mad);
}

}  // namespace

void launch_trim(const float *raw, const int64_t *off, const int *nsample, int nread, int chunk, float perc,
                 int trim_start, int trim_end, float *mads, const int64_t *mads_off, int *start_end, cudaStream_t s) {
    if (nread <= 0) return;
    trim_kernel<<<nread, PREP_THREADS, 0, s>>>(raw, off, nsample, chunk, perc, trim_start, trim_end, mads, mads_off,
                                               reinterpret_cast<int2 *>(start_end));
}

void launch_medmad(const float *src, const int64_t *src_off, float *dst, const int64_t *dst_off, const int *nsample,
                   int nread, cudaStream_t s) {
    if (nread <= 0) return;
    medmad_kernel<<<nread, PREP_THREADS, 0, s>>>(src, src_off, dst, dst_off, nsample);
}

}  // namespace sb2
