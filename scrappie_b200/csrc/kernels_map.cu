// map_to_sequence_viterbi / _forward / _viterbi_banded / _forward_banded (src/decode.c:1420-1964):
// local-global alignment of a log posterior to a known k-mer sequence (python `map_post_to_sequence`, seqmappy).
//
// One CTA per alignment; thread t owns sequence positions t, t + 256, ...  The two score vectors of the reference
// (seqlen + 2 floats each: positions, start state, end state) alternate in global scratch (L1/L2 resident); one
// __syncthreads per block.  Every sum is evaluated in the reference's order and only strict improvements replace
// a score, so the Viterbi variants are bit-exact (score and path); the forward variants use
// logsumexpf = fmaxf + log1pf(expf(-|x - y|)) like the reference.  The banded variants rewrite only the positions
// inside a block's band, exactly as the reference does -- positions outside keep their value from two blocks
// earlier.  Traceback: one byte per (block, position) -- 0 stay, 1 step, 2 skip, 3 from the start state.
#include <math.h>

#include "kernels.h"

namespace sb2 {
namespace {

constexpr int MAP_THREADS = 256;
constexpr float MAP_BIG = 1.e30f;

__device__ __forceinline__ float lse2f(float x, float y) { return fmaxf(x, y) + log1pf(expf(-fabsf(x - y))); }

template <bool FWD>
__device__ __forceinline__ float comb(float a, float b) { return FWD ? lse2f(a, b) : fmaxf(a, b); }

template <bool FWD, bool BANDED>
__global__ void __launch_bounds__(MAP_THREADS)
map_to_sequence_kernel(const float *__restrict__ lp, int nblock, int nst, int stride, float stay_pen, float skip_pen,
                       float local_pen, const int *__restrict__ seq, int seqlen, const int *__restrict__ low,
                       const int *__restrict__ high, float *buf, uint8_t *tb, uint8_t *tb_end, float *score_out,
                       int *path) {
    const int tid = threadIdx.x;
    const int STAY = nst - 1, START = seqlen, END = seqlen + 1, ns = seqlen + 2;
    float *c = buf, *p = buf + ns;
    for (int i = tid; i < ns; i += MAP_THREADS) { c[i] = -MAP_BIG; p[i] = -MAP_BIG; }
    __syncthreads();
    if (tid == 0) { if (BANDED) p[START] = 0.0f; else c[START] = 0.0f; }
    __syncthreads();

    for (int blk = 0; blk < nblock; blk++) {
        const float *l = lp + (size_t)blk * stride;
        if (!(BANDED && blk == 0)) { float *t = p; p = c; c = t; }
        const float lstay = l[STAY];
        if (tid == 0) {
            const float loc = FWD ? lse2f(-local_pen, lstay) : fmaxf(-local_pen, lstay);
            c[START] = p[START] + loc;
            float e = p[END] + loc;
            uint8_t from_seq = 0;
            if (BANDED && blk == 0) e = comb<FWD>(e, p[START] - local_pen);
            const float se = p[seqlen - 1] - local_pen;
            if (FWD || BANDED) e = comb<FWD>(e, se);
            else if (se > e) { e = se; from_seq = 1; }
            c[END] = e;
            if (tb_end) tb_end[blk] = from_seq;
        }
        if (!BANDED) {
            uint8_t *t = tb ? tb + (size_t)blk * seqlen : nullptr;
            for (int pos = tid; pos < seqlen; pos += MAP_THREADS) {
                float s = (p[pos] - stay_pen) + lstay;
                uint8_t code = 0;
                const float lk = l[seq[pos]];
                if (pos >= 1) {
                    const float sc = p[pos - 1] + lk;
                    if (FWD) s = lse2f(s, sc);
                    else if (sc > s) { s = sc; code = 1; }
                }
                if (pos >= 2) {
                    const float sc = (p[pos - 2] - skip_pen) + lk;
                    if (FWD) s = lse2f(s, sc);
                    else if (sc > s) { s = sc; code = 2; }
                }
                if (pos == 0) {
                    const float sc = p[START] + lk;
                    if (FWD) s = lse2f(s, sc);
                    else if (sc > s) { s = sc; code = 3; }
                }
                c[pos] = s;
                if (t) t[pos] = code;
            }
        } else if (blk == 0) {
            if (tid == 0) {
                float c0 = comb<FWD>(c[0], (p[0] + lstay) - stay_pen);
                if (high[0] > 0) c[1] = l[seq[1]];
                if (high[0] > 1) c[2] = l[seq[2]] - skip_pen;
                c0 = comb<FWD>(c0, p[START] + l[seq[0]]);
                c[0] = c0;
            }
        } else {
            const int lo = low[blk], lo1 = low[blk - 1], hi = high[blk], hi1 = high[blk - 1];
            const int a1 = max(lo, lo1 + 1), b1 = min(hi, hi1 + 1), a2 = max(lo, lo1 + 2), b2 = min(hi, hi1 + 2);
            const int first = min(lo, min(a1, a2)), last = max(hi1, max(b1, b2));
            for (int pos = first + tid; pos < last; pos += MAP_THREADS) {
                float s = c[pos];
                bool touched = false;
                if (pos >= lo && pos < hi1) { s = (p[pos] - stay_pen) + lstay; touched = true; }
                if (pos >= a1 && pos < b1) { s = comb<FWD>(p[pos - 1] + l[seq[pos]], s); touched = true; }
                if (pos >= a2 && pos < b2) { s = comb<FWD>((p[pos - 2] - skip_pen) + l[seq[pos]], s); touched = true; }
                if (touched) c[pos] = s;
            }
            // move from the start state into the sequence: position 0 belongs to thread 0 whenever lo == 0
            if (lo == 0 && tid == 0) c[0] = comb<FWD>(c[0], p[START] + l[seq[0]]);
        }
        __syncthreads();
    }

    if (tid == 0) {
        const float a = c[seqlen - 1], e = c[END];
        *score_out = comb<FWD>(a, e);
        if (path != nullptr && tb != nullptr) {
            int cur = (a > e) ? seqlen - 1 : END;
            path[nblock - 1] = cur;
            for (int blk = nblock - 1; blk > 0; blk--) {
                if (cur == END) cur = tb_end[blk] ? seqlen - 1 : END;
                else if (cur != START) {
                    const uint8_t code = tb[(size_t)blk * seqlen + cur];
                    cur = (code == 3) ? START : cur - (int)code;
                }
                path[blk - 1] = cur;
            }
            for (int blk = 0; blk < nblock; blk++)
                if (path[blk] == START || path[blk] == END) path[blk] = -1;
        }
    }
}

}  // namespace

void launch_map_to_sequence(const float *lp, int nblock, int nst, int stride, float stay_pen, float skip_pen,
                            float local_pen, const int *seq, int seqlen, const int *low, const int *high, int forward,
                            float *buf, uint8_t *tb, uint8_t *tb_end, float *score, int *path, cudaStream_t s) {
    const bool banded = (low != nullptr && high != nullptr);
#define SB2_MAP(F, B) map_to_sequence_kernel<F, B><<<1, MAP_THREADS, 0, s>>>(lp, nblock, nst, stride, stay_pen, skip_pen, \
                          local_pen, seq, seqlen, low, high, buf, tb, tb_end, score, path)
    if (forward) { if (banded) SB2_MAP(true, true); else SB2_MAP(true, false); }
    else { if (banded) SB2_MAP(false, true); else SB2_MAP(false, false); }
#undef SB2_MAP
}

}  // namespace sb2
