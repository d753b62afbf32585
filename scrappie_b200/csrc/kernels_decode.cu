// Transducer Viterbi, fourth generation: ONE WARP PER READ, all 1024 k-mer scores in registers.
//
// Replaces decode_transducer + viterbi_local_backtrace (src/decode.c:123-365, :58-98) for the 1024-history
// models without slip; results are bit-identical to the earlier generations (tests/test_gpu_parity.py).
//
// Why: the CTA-per-read kernels (kernels_v1.cu) pay two __syncthreads and a shared-memory round trip of the
// whole score vector per block and need 256 threads x 48 registers per read.  Here lane L of the warp owns
// the 32 states  s = 128 g + 4 L + j  (g = 0..7, j = 0..3):
//   * the log-posterior columns stream through a per-warp shared-memory ring filled by cp.async (eight
//     512-byte-contiguous 16-byte copies per column and lane), NS - 1 columns ahead of the one in use, so no
//     register is tied up by loads in flight and the HBM latency is off the per-block dependency chain,
//   * the 4-way "step" maximum  m4[t] = max_q prev[256 q + t]  is LOCAL to a lane: for t = 128 g0 + 4 L + j0
//     the four predecessors are this lane's registers (g = 2 q + g0, j = j0),
//   * the 16-way "skip" maximum m16[u] = max_b m4[u + 64 b] needs one exchange with lane L ^ 16,
//   * the maxima are handed to the states that consume them (state s reads m4[s >> 2], m16[s >> 4]: one
//     LDS.64 each per FOUR states) through a 5 KB per-warp shared-memory table, double-buffered so that a
//     single __syncwarp per block suffices,
//   * traceback codes are written as two 512-byte-contiguous STG.128 per block; the backtrace reads them
//     32 blocks at a time (one lane per block, speculating that the path stays) instead of staging whole rows.
// A warp needs no block-level barrier and ~5 KB of shared memory, so decode CTAs co-reside with the GRU scan
// CTAs of other batches instead of waiting for free SMs.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace sb2 {
namespace {

constexpr float DEC_BIG = 1.e30f;
enum { TB_STAY = 0, TB_STEP = 1, TB_SKIP = 5, TB_START = 85 };       // same codes as kernels_v1.cu
constexpr int NH = 1024;
constexpr int NS = 4;               // depth of the cp.async column ring
constexpr int RING_ROW = NH + 32;   // floats per ring slot: 1024 k-mer log-posteriors + one copy of the stay row per lane

struct WarpTables {
    float2 m4[2][256];              // (step maximum, traceback code TB_STEP + q) per suffix t
    float2 m16[2][64];              // (skip maximum, traceback code TB_SKIP + r16) per suffix u
};

__device__ __forceinline__ float redux_max(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
__device__ __forceinline__ int redux_min(int v) {
    int m;
    asm volatile("redux.sync.min.s32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(v));
    return m;
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Traceback codes travel through the update as floats whose bit pattern is CODE_BIAS | code (2^23 + code: exact),
// so that "if (s < cand) { s = cand; code = c; }" -- strict improvement only, like the reference's masked max -- can
// be ONE compare on the half-rate ALU pipe plus two predicated FFMAs (x * 1.0f + -0.0f == x exactly) on the FMA
// pipe.  Written as a C conditional ptxas emits FSETP + FSEL + SEL, three ALU-pipe instructions, and that pipe
// bounds the kernel (profiles/r18); the multiplier and addend are kernel arguments so they cannot be folded away.
constexpr uint32_t CODE_BIAS = 0x4B000000u;
__device__ __forceinline__ float code_f(uint32_t code) { return __uint_as_float(CODE_BIAS | code); }
template <bool FMA>
__device__ __forceinline__ void take_if_better(float &s, float &code, float cand, float c, float one, float negzero) {
    if (FMA) {
        asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %0, %2;\n\t@p fma.rn.f32 %0, %2, %4, %5;\n\t@p fma.rn.f32 %1, %3, %4, %5;\n\t}"
            : "+f"(s), "+f"(code) : "f"(cand), "f"(c), "f"(one), "f"(negzero));
    } else {
        if (s < cand) { s = cand; code = c; }
    }
}

// byte position of state s inside a 1024-byte traceback row
__device__ __forceinline__ int tb_offset(int s) {
    const int g = s >> 7, L = (s >> 2) & 31, j = s & 3;
    return (g >> 2) * 512 + L * 16 + (g & 3) * 4 + j;
}

template <int WPC, bool FMA>
__global__ void __launch_bounds__(32 * WPC)
decode_transducer_warp_kernel(const float *__restrict__ post, BatchDims d, int ostride, float stay_pen,
                              float skip_pen, float local_pen, uint8_t *tb, int *tb_end, int *path,
                              float *__restrict__ score, float one, float negzero) {
    __shared__ __align__(16) WarpTables tables[WPC];
    __shared__ __align__(16) float rings[WPC][NS * RING_ROW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * WPC + warp;
    if (r >= d.nread) return;                       // warps are independent: no block-level barrier below
    WarpTables &ws = tables[warp];
    const int T = d.nblock[r];
    const float *lp = post + (size_t)d.col_off[r] * ostride;
    uint8_t *tbr = tb + (size_t)d.col_off[r] * NH;
    int *tbe = tb_end + d.col_off[r];
    constexpr int BIG_IDX = 0x7fffffff;
    const int hbit = lane >> 4;

    float cur[8][4];
#pragma unroll
    for (int g = 0; g < 8; g++)
#pragma unroll
        for (int j = 0; j < 4; j++) cur[g][j] = -DEC_BIG;
    float curS = 0.0f, curE = -DEC_BIG;             // start / end state, replicated in every lane
    // column ring: slot (c % NS) holds column c; this lane copies and later reads only its own 32 + 1 values
    float *ring = rings[warp];
    auto fetch_column = [&](int c) {
        if (c < T) {
            const float *src = lp + (size_t)c * ostride;
            float *dst = ring + (c % NS) * RING_ROW;
#pragma unroll
            for (int g = 0; g < 8; g++) cp_async16(dst + 128 * g + 4 * lane, src + 128 * g + 4 * lane);
            cp_async4(dst + NH + lane, src + NH);
        }
        cp_async_commit();                          // one group per column, empty past the end
    };
#pragma unroll
    for (int c = 0; c < NS - 1; c++) fetch_column(c);
    uint8_t *tb_cur = tbr + lane * 16;

    for (int blk = 0; blk < T; blk++, tb_cur += NH) {
        const int buf = blk & 1;
        fetch_column(blk + NS - 1);                 // into the slot column blk - 1 was read from
        cp_async_wait<NS - 1>();                    // column blk has landed
        const float *colv = ring + (blk % NS) * RING_ROW;
        const float stay = colv[NH + lane] - stay_pen;

        // ---- step maxima of this lane's eight suffixes t = 128 g0 + 4 lane + j0 (ascending scan, strict <)
        float m4v[2][4];
        int m4r[2][4];
#pragma unroll
        for (int g0 = 0; g0 < 2; g0++)
#pragma unroll
            for (int j0 = 0; j0 < 4; j0++) {
                float b = cur[g0][j0];
                float q4 = code_f(0);
#pragma unroll
                for (int q = 1; q < 4; q++) take_if_better<FMA>(b, q4, cur[2 * q + g0][j0], code_f((uint32_t)q), one, negzero);
                m4v[g0][j0] = b;
                m4r[g0][j0] = (int)(__float_as_uint(q4) & 3u);
            }
#pragma unroll
        for (int g0 = 0; g0 < 2; g0++) {
            float4 *dst = reinterpret_cast<float4 *>(&ws.m4[buf][128 * g0 + 4 * lane]);
            dst[0] = make_float4(m4v[g0][0], code_f(TB_STEP + m4r[g0][0]), m4v[g0][1], code_f(TB_STEP + m4r[g0][1]));
            dst[1] = make_float4(m4v[g0][2], code_f(TB_STEP + m4r[g0][2]), m4v[g0][3], code_f(TB_STEP + m4r[g0][3]));
        }

        // ---- skip maxima: suffix u = 4 (lane & 15) + j0 combines m4[u + 64 b], b = 2 g0 + (lane >> 4);
        //      r16 = 4 q + b, the lowest among the maximal entries (reference: ascending scan over r16)
        float m16v[4];
        int m16c[4];
#pragma unroll
        for (int j0 = 0; j0 < 4; j0++) {
            const float v0 = m4v[0][j0], v1 = m4v[1][j0];
            const int i0 = 4 * m4r[0][j0] + hbit, i1 = 4 * m4r[1][j0] + 2 + hbit;
            const bool take1 = (v1 > v0) || (v1 == v0 && i1 < i0);
            const float pv = take1 ? v1 : v0;
            const int pi = take1 ? i1 : i0;
            const float ov = __shfl_xor_sync(0xffffffffu, pv, 16);
            const int oi = __shfl_xor_sync(0xffffffffu, pi, 16);
            const bool takeo = (ov > pv) || (ov == pv && oi < pi);
            m16v[j0] = takeo ? ov : pv;
            m16c[j0] = TB_SKIP + (takeo ? oi : pi);
        }
        if (lane < 16) {
            float4 *dst = reinterpret_cast<float4 *>(&ws.m16[buf][4 * lane]);
            dst[0] = make_float4(m16v[0], code_f(m16c[0]), m16v[1], code_f(m16c[1]));
            dst[1] = make_float4(m16v[2], code_f(m16c[2]), m16v[3], code_f(m16c[3]));
        }

        // ---- end state (src/decode.c:345-356): best of "stay in end" and max_s (prev[s] - local_pen), the
        //      lowest s among the maximal ROUNDED candidates.  Rounding is monotone, so the maximum is
        //      fl(M - local_pen); unless the next float below M rounds to the same value (rare) only states
        //      with prev == M qualify and the lane-local hierarchy gives the lowest of them.
        {
            float ml = fmaxf(fmaxf(fmaxf(m4v[0][0], m4v[0][1]), fmaxf(m4v[0][2], m4v[0][3])),
                             fmaxf(fmaxf(m4v[1][0], m4v[1][1]), fmaxf(m4v[1][2], m4v[1][3])));
            const float M = redux_max(ml);
            const float wv = M - local_pen;
            float e = curE + fmaxf(-local_pen, stay);
            int from = NH + 1;
            if (wv > e) {                           // warp-uniform
                const float below = __int_as_float(__float_as_int(M) + ((M > 0.0f) ? -1 : 1));
                int cand = BIG_IDX;
                if (M != 0.0f && (below - local_pen) != wv) {
#pragma unroll
                    for (int g0 = 1; g0 >= 0; g0--)
#pragma unroll
                        for (int j0 = 3; j0 >= 0; j0--) {
                            const int idx = 256 * m4r[g0][j0] + 128 * g0 + 4 * lane + j0;
                            cand = (m4v[g0][j0] == M) ? min(cand, idx) : cand;
                        }
                } else {
#pragma unroll
                    for (int g = 0; g < 8; g++)
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            cand = ((cur[g][j] - local_pen) == wv) ? min(cand, 128 * g + 4 * lane + j) : cand;
                }
                from = redux_min(cand);
                e = wv;
            }
            if (lane == 0) tbe[blk] = from;
            curE = e;
        }
        __syncwarp();

        // ---- state updates in the reference's order: stay, step, skip, from-start (strict improvements only)
        uint32_t codes[8];
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const float2 e4 = ws.m4[buf][32 * g + lane];
            const float2 e16 = ws.m16[buf][8 * g + (lane >> 2)];
            const float4 l4 = *reinterpret_cast<const float4 *>(colv + 128 * g + 4 * lane);
            const float lpj[4] = {l4.x, l4.y, l4.z, l4.w};
            float cj[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float s = cur[g][j] + stay;
                float code = code_f(TB_STAY);
                take_if_better<FMA>(s, code, lpj[j] + e4.x, e4.y, one, negzero);
                take_if_better<FMA>(s, code, (lpj[j] + e16.x) - skip_pen, e16.y, one, negzero);
                take_if_better<FMA>(s, code, curS + lpj[j], code_f(TB_START), one, negzero);
                cur[g][j] = s;
                cj[j] = code;
            }
            // low bytes of the four biased codes -> one 32-bit word
            const uint32_t lo2 = __byte_perm(__float_as_uint(cj[0]), __float_as_uint(cj[1]), 0x0040);
            const uint32_t hi2 = __byte_perm(__float_as_uint(cj[2]), __float_as_uint(cj[3]), 0x0040);
            const uint32_t cw = __byte_perm(lo2, hi2, 0x5410);
            codes[g] = cw;
        }
        *reinterpret_cast<uint4 *>(tb_cur) = make_uint4(codes[0], codes[1], codes[2], codes[3]);
        *reinterpret_cast<uint4 *>(tb_cur + 512) = make_uint4(codes[4], codes[5], codes[6], codes[7]);
        curS = curS + fmaxf(-local_pen, stay);
    }

    // ---- final argmax over (states..., start, end): first maximum wins (argmaxf, src/util.c:9-23)
    int last;
    {
        float ml = -INFINITY;
        int cand = BIG_IDX;
#pragma unroll
        for (int g = 0; g < 8; g++)
#pragma unroll
            for (int j = 0; j < 4; j++) ml = fmaxf(ml, cur[g][j]);
        float bv = redux_max(ml);
#pragma unroll
        for (int g = 7; g >= 0; g--)
#pragma unroll
            for (int j = 3; j >= 0; j--) cand = (cur[g][j] == bv) ? (128 * g + 4 * lane + j) : cand;
        last = redux_min(cand);
        if (curS > bv) { bv = curS; last = NH; }
        if (curE > bv) { bv = curE; last = NH + 1; }
        if (lane == 0) score[r] = bv;
    }
    __syncwarp();                                   // traceback rows written by the other lanes are visible

    // ---- backtrace, 32 blocks per round: lane k looks at block blk - k under the assumption that the
    //      path has not moved; the first lane that sees a move ends the round
    int *seq = path + d.col_off[r] + r;
    int blk = T - 1;
    while (blk >= 0) {
        const int myrow = blk - lane;
        if (last == NH) {                           // the start state only ever follows itself
            for (int i = lane; i <= blk; i += 32) seq[i + 1] = NH;
            blk = -1;
        } else if (last == NH + 1) {
            const int v = (myrow >= 0) ? tbe[myrow] : -2;
            const unsigned m = __ballot_sync(0xffffffffu, v != NH + 1);
            const int f = m ? (__ffs(m) - 1) : 32;
            if (myrow >= 0 && lane <= f) seq[myrow + 1] = NH + 1;
            if (f < 32 && blk - f >= 0) {
                last = __shfl_sync(0xffffffffu, v, f);
                blk -= f + 1;
            } else {
                blk = (f < 32) ? -1 : blk - 32;
            }
        } else {
            const int code = (myrow >= 0) ? (int)tbr[(size_t)myrow * NH + tb_offset(last)] : 255;
            const unsigned m = __ballot_sync(0xffffffffu, code != TB_STAY);
            const int f = m ? (__ffs(m) - 1) : 32;
            if (myrow >= 0) {
                if (lane < f) seq[myrow + 1] = -1;
                else if (lane == f) seq[myrow + 1] = last;
            }
            if (f < 32 && blk - f >= 0) {
                const int c = __shfl_sync(0xffffffffu, code, f);
                if (c >= TB_START) last = NH;
                else if (c >= TB_SKIP) last = (c - TB_SKIP) * (NH / 16) + last / 16;
                else last = (c - TB_STEP) * (NH / 4) + last / 4;
                blk -= f + 1;
            } else {
                blk = (f < 32) ? -1 : blk - 32;
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        seq[0] = last;
        for (int i = 0; i < T; i++) { if (seq[i] == NH) seq[i] = -1; else break; }
        for (int i = T; i >= 0; i--) { if (seq[i] == NH + 1) seq[i] = -1; else break; }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------
// posterior_crf (src/decode.c:928-1012): forward-backward over the 5 x 5 transition energies of rnnrf_r94
// ---------------------------------------------------------------------------------
// One warp per read.  Lanes 0..24 fetch the 25 energies of a block with one coalesced load; lane s < 5 owns
// state s: the forward value alpha[s] (written to the output column as the reference does, then read back by
// the same lane in the backward sweep) and the backward value beta[s].  Every logsumexp fold runs in the
// reference's order (ascending source state) with logsumexpf = fmaxf + log1pf(expf(-|x - y|)); the column
// normaliser starts its fold at 0.0f exactly like the reference (its probabilities sum to S / (1 + S)).
namespace {

constexpr int PCRF_WARPS = 4;

__device__ __forceinline__ float lse2(float x, float y) { return fmaxf(x, y) + log1pf(expf(-fabsf(x - y))); }

__global__ void __launch_bounds__(32 * PCRF_WARPS)
posterior_crf_kernel(const float *__restrict__ trans, BatchDims d, int ostride, float *__restrict__ post) {
    constexpr int NS = 5, PS = 8;                   // states, output stride (4 * ceil(5 / 4))
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * PCRF_WARPS + warp;
    if (r >= d.nread) return;
    const int T = d.nblock[r];
    const float *tr = trans + (size_t)d.col_off[r] * ostride;
    float *out = post + ((size_t)d.col_off[r] + r) * PS;            // read r owns nblock + 1 columns
    const bool own = lane < NS;
    const int me = own ? lane : 0;
    const unsigned FULL = 0xffffffffu;

    // ---- forward sweep: alpha[0] = 0, alpha[b + 1][to] = LSE_from (trans[b][to][from] + alpha[b][from])
    float a = 0.0f;
    if (lane < PS) out[lane] = 0.0f;
    float tv_next = (T > 0 && lane < NS * NS) ? tr[lane] : 0.0f;
    for (int blk = 0; blk < T; blk++) {
        const float tv = tv_next;
        if (blk + 1 < T && lane < NS * NS) tv_next = tr[(size_t)(blk + 1) * ostride + lane];
        float acc = 0.0f;
#pragma unroll
        for (int from = 0; from < NS; from++) {
            const float v = __shfl_sync(FULL, tv, me * NS + from) + __shfl_sync(FULL, a, from);
            acc = (from == 0) ? v : lse2(acc, v);
        }
        a = acc;
        if (lane < PS) out[(size_t)(blk + 1) * PS + lane] = own ? a : 0.0f;
    }

    // ---- last column: normalise alpha[T]
    {
        float tot = 0.0f;
#pragma unroll
        for (int st = 0; st < NS; st++) tot = lse2(tot, __shfl_sync(FULL, a, st));
        if (own) out[(size_t)T * PS + lane] = expf(a - tot);
    }

    // ---- backward sweep: beta[T] = 0, beta[b][from] = LSE_to (trans[b][to][from] + beta[b + 1][to]);
    //      column b becomes exp(alpha[b] + beta[b] - normaliser)
    float b = 0.0f;
    float tvb_next = (T > 0 && lane < NS * NS) ? tr[(size_t)(T - 1) * ostride + lane] : 0.0f;
    float al_next = (T > 0 && own) ? out[(size_t)(T - 1) * PS + lane] : 0.0f;
    for (int blk = T; blk > 0; blk--) {
        const float tv = tvb_next, al = al_next;
        if (blk > 1) {
            if (lane < NS * NS) tvb_next = tr[(size_t)(blk - 2) * ostride + lane];
            if (own) al_next = out[(size_t)(blk - 2) * PS + lane];
        }
        float acc = 0.0f;
#pragma unroll
        for (int to = 0; to < NS; to++) {
            const float v = __shfl_sync(FULL, tv, to * NS + me) + __shfl_sync(FULL, b, to);
            acc = (to == 0) ? v : lse2(acc, v);
        }
        b = acc;
        const float c = al + b;
        float tot = 0.0f;
#pragma unroll
        for (int st = 0; st < NS; st++) tot = lse2(tot, __shfl_sync(FULL, c, st));
        if (own) out[(size_t)(blk - 1) * PS + lane] = expf(c - tot);
    }
}

}  // namespace

void launch_posterior_crf(const float *trans, const BatchDims &d, int ostride, float *post, cudaStream_t s) {
    posterior_crf_kernel<<<(d.nread + PCRF_WARPS - 1) / PCRF_WARPS, 32 * PCRF_WARPS, 0, s>>>(trans, d, ostride, post);
}

void launch_decode_transducer_warp(const float *post, const BatchDims &d, int ostride, float stay_pen,
                                   float skip_pen, float local_pen, uint8_t *tb, int *tb_end, int *path,
                                   float *score, cudaStream_t s) {
    // read once (thread-safe static initialisation): warps per CTA, and "sel" = selects on the ALU pipe (cross-check)
    static const int wpc = [] { const char *e = getenv("SCRAPPIE_B200_DECODE_WPC"); return e ? atoi(e) : 2; }();
    static const int fma = [] { const char *e = getenv("SCRAPPIE_B200_DECODE_SEL"); return (e && 0 == strcmp(e, "sel")) ? 0 : 1; }();
#define SB2_LAUNCH_WARP_DECODE(WPC, FMA)                                                                  \
    decode_transducer_warp_kernel<WPC, FMA><<<(d.nread + WPC - 1) / WPC, 32 * WPC, 0, s>>>(                \
        post, d, ostride, stay_pen, skip_pen, local_pen, tb, tb_end, path, score, 1.0f, -0.0f)
    if (wpc == 1) { if (fma) SB2_LAUNCH_WARP_DECODE(1, true); else SB2_LAUNCH_WARP_DECODE(1, false); }
    else { if (fma) SB2_LAUNCH_WARP_DECODE(2, true); else SB2_LAUNCH_WARP_DECODE(2, false); }
#undef SB2_LAUNCH_WARP_DECODE
}

}  // namespace sb2
